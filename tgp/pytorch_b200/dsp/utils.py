"""Helpers with the reference's names (code/dsp/utils.py:39-55, 143-159)."""
import torch

from . import config as cg
from ..functional import NanError, NumericalWarning  # noqa: F401  (re-exported, as the reference imports them here)


def positive_transform(x):
    if cg.positive_transform == 'exp':
        return torch.exp(x)
    if cg.positive_transform == 'softplus':
        return torch.log(torch.exp(x) + 1)
    raise NotImplementedError('positive_transform %r' % (cg.positive_transform,))


def inverse_positive_transform(x):
    if cg.positive_transform == 'exp':
        return torch.log(x)
    if cg.positive_transform == 'softplus':
        return torch.log(torch.exp(x) - 1.)
    raise NotImplementedError('inverse_positive_transform %r' % (cg.positive_transform,))


def inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


def KMEANS(X, num_Z, n_init=1, seed=None):
    """Inducing-point initialisation, caller-side and CPU (sklearn), as in the reference."""
    from sklearn.cluster import KMeans
    if seed is None:
        seed = cg.config_seed
    km = KMeans(n_clusters=num_Z, init='k-means++', n_init=n_init, random_state=seed).fit(X.to('cpu').numpy())
    return torch.tensor(km.cluster_centers_, dtype=cg.dtype).to(cg.device)
