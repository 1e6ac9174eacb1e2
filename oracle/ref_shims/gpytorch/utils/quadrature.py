import math
import numpy as np
import torch
from .broadcasting import _pad_with_singletons


class GaussHermiteQuadrature1D(torch.nn.Module):
    """hermgauss(n) nodes/weights cast with torch.Tensor() (default dtype at construction)."""

    def __init__(self, num_locs=20):
        super().__init__()
        self.num_locs = num_locs
        locs, wts = np.polynomial.hermite.hermgauss(num_locs)
        self.locations = torch.Tensor(locs)
        self.weights = torch.Tensor(wts)

    def _apply(self, fn):
        self.locations = fn(self.locations)
        self.weights = fn(self.weights)
        return super()._apply(fn)

    def forward(self, func, gaussian_dists):
        mean = gaussian_dists.mean
        var = gaussian_dists.variance
        locs = _pad_with_singletons(self.locations, 0, mean.dim())
        shifted = torch.sqrt(2.0 * var) * locs + mean
        vals = func(shifted)
        wts = _pad_with_singletons(self.weights, 0, vals.dim() - 1)
        res = (1 / math.sqrt(math.pi)) * (vals * wts)
        return res.sum(tuple(range(self.locations.dim())))
