Point = Polygon = MultiPolygon = object
