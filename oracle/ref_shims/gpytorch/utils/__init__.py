from . import broadcasting, quadrature, transforms, errors, warnings, cholesky  # noqa: F401
