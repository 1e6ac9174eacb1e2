"""RBF / Scale kernels as gpytorch 1.1.1 evaluates them."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from .lazy import NonLazyTensor, delazify


class Kernel(nn.Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self._batch_shape = batch_shape
        self.ard_num_dims = ard_num_dims
        if self.has_lengthscale:
            nd = 1 if ard_num_dims is None else ard_num_dims
            self.raw_lengthscale = nn.Parameter(torch.zeros(*batch_shape, 1, nd))

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        return F.softplus(self.raw_lengthscale)

    def __call__(self, x1, x2=None, diag=False, last_dim_is_batch=False, **params):
        if x2 is None:
            x2 = x1
        out = self.forward(x1, x2, diag=diag, **params)
        if diag:
            return delazify(out)
        return out if hasattr(out, 'evaluate') else NonLazyTensor(out)


def sq_dist(x1, x2, same):
    # quadratic expansion after centring on x1's mean; diagonal pinned to 0 only for identical,
    # grad-free operands; negatives clamped
    shift = x1.mean(-2, keepdim=True)
    x1 = x1 - shift
    x2 = x2 - shift
    n1 = x1.pow(2).sum(-1, keepdim=True)
    o1 = torch.ones_like(n1)
    pin = same and not x1.requires_grad and not x2.requires_grad
    if pin:
        n2, o2 = n1, o1
    else:
        n2 = x2.pow(2).sum(-1, keepdim=True)
        o2 = torch.ones_like(n2)
    lhs = torch.cat([-2.0 * x1, n1, o1], dim=-1)
    rhs = torch.cat([x2, o2, n2], dim=-1)
    res = lhs.matmul(rhs.transpose(-2, -1))
    if pin:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min_(0)


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, diag=False, **params):
        a = x1.div(self.lengthscale)
        b = x2.div(self.lengthscale)
        same = torch.equal(a, b)
        if diag:
            if same:
                return torch.zeros(*a.shape[:-1], dtype=x1.dtype, device=x1.device).div_(-2).exp_()
            return (a - b).norm(p=2, dim=-1).pow(2).div(-2).exp()
        return sq_dist(a, b, same).div(-2).exp()


class MaternKernel(RBFKernel):
    def __init__(self, nu=2.5, **kwargs):
        super().__init__(**kwargs)
        self.nu = nu

    def forward(self, *a, **k):
        raise NotImplementedError('Matern is outside the hot-path scope (SURVEY.md §2.1 row 11)')


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), **kwargs):
        super().__init__(batch_shape=batch_shape)
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(*batch_shape))

    @property
    def outputscale(self):
        return F.softplus(self.raw_outputscale)

    def forward(self, x1, x2, diag=False, **params):
        base = delazify(self.base_kernel.forward(x1, x2, diag=diag, **params))
        s = self.outputscale
        return base * (s.unsqueeze(-1) if diag else s.view(*s.shape, 1, 1))


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)


class ProductKernel(AdditiveKernel):
    pass
