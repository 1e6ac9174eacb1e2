def _unused(*a, **k):
    pass
