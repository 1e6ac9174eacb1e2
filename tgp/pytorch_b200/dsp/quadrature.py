"""Gauss-Hermite rule holder with the attribute surface of gpytorch's GaussHermiteQuadrature1D
(`locations`, `weights`, callable on (func, Normal)); the fused kernels take the same nodes as device arrays."""
import math

import numpy as np
import torch


class GaussHermiteQuadrature1D(torch.nn.Module):
    def __init__(self, num_locs=20):
        super().__init__()
        self.num_locs = num_locs
        locs, wts = np.polynomial.hermite.hermgauss(num_locs)
        self.locations = torch.tensor(locs, dtype=torch.get_default_dtype())
        self.weights = torch.tensor(wts, dtype=torch.get_default_dtype())

    def _apply(self, fn):
        self.locations = fn(self.locations)
        self.weights = fn(self.weights)
        return super()._apply(fn)

    def forward(self, func, gaussian_dists):
        """Generic (un-fused) rule for arbitrary integrands; not used by the ELBO / test-NLL path."""
        mean, var = gaussian_dists.mean, gaussian_dists.variance
        shape = [-1] + [1] * mean.dim()
        vals = func(torch.sqrt(2.0 * var) * self.locations.view(shape) + mean)
        wshape = [-1] + [1] * (vals.dim() - 1)
        return ((1 / math.sqrt(math.pi)) * (vals * self.weights.view(wshape))).sum(0)
