nearest_points = None
