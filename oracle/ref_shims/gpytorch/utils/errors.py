class NanError(RuntimeError):
    pass
