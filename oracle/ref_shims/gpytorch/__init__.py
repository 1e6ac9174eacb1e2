"""Minimal stand-in for gpytorch 1.1.1 (only the surface the reference hot path uses)."""
__version__ = '1.1.1'
from . import utils, lazy, means, kernels, variational  # noqa: F401
