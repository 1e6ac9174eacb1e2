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


def KMEANS_device(X, num_Z, n_init=1, seed=None, max_iter=100, tol=1e-4):
    """Inducing-point initialisation by Lloyd's k-means ON THE DEVICE (tgp_kmeans_iteration) — for data sets where the
    reference's host-side sklearn call (`KMEANS`, n_init=10 in main.py:145) takes hours (N = 5 M, M = 1024).  Restarts are
    seeded greedy k-means++ initialisations (D^2 sampling with 2+ln(k) candidates per centre, the same rule sklearn uses; a few
    element-wise device passes per centre); the restart with the lowest inertia wins; iteration stops when the inertia improves
    by less than `tol` relative.  Not bit-comparable with sklearn (different random stream and arithmetic order);
    `tests/test_gpu_kmeans.py` checks the fixed-point property and the inertia against sklearn's."""
    import ctypes as C
    from .. import _lib
    lib = _lib.load()
    if not X.is_cuda:
        raise RuntimeError('KMEANS_device clusters a CUDA tensor (use KMEANS for the host-side sklearn path)')
    Xd = X.detach().to(torch.float64).contiguous()
    N, D = Xd.shape
    if num_Z > N:
        raise ValueError('more inducing points than rows')
    dev = Xd.device
    gen = torch.Generator(device=dev).manual_seed(cg.config_seed if seed is None else int(seed))

    def plus_plus():
        first = int(torch.randint(0, N, (1,), generator=gen, device=dev))
        idx = [first]
        mind = ((Xd - Xd[first]) ** 2).sum(1)
        import math
        trials = 2 + int(math.log(num_Z))                        # greedy k-means++, as sklearn: several candidates per centre,
        for _j in range(1, num_Z):                               # keep the one that lowers the potential most
            cand = torch.multinomial(mind.clamp_min(0) + 1e-300, trials, replacement=True, generator=gen)
            d = ((Xd[:, None, :] - Xd[cand][None, :, :]) ** 2).sum(2) if N * trials * D <= (1 << 26) else \
                torch.stack([((Xd - Xd[c]) ** 2).sum(1) for c in cand.tolist()], dim=1)
            pot = torch.minimum(mind[:, None], d)
            b = int(pot.sum(0).argmin())
            idx.append(int(cand[b]))
            mind = pot[:, b].contiguous()
        return Xd[torch.tensor(idx, device=dev)].clone()

    sums = torch.empty(num_Z, D, dtype=torch.float64, device=dev)
    counts = torch.empty(num_Z, dtype=torch.float64, device=dev)
    inertia = torch.empty(1, dtype=torch.float64, device=dev)
    best, best_inertia = None, float('inf')
    with torch.cuda.device(dev):
        for _ in range(max(1, n_init)):
            Cn = plus_plus()
            prev = float('inf')
            for it in range(max_iter + 1):
                st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(lib.tgp_kmeans_iteration(Xd.data_ptr(), N, D, Cn.data_ptr(), num_Z, None, sums.data_ptr(), counts.data_ptr(),
                                                    inertia.data_ptr(), 1 if it < max_iter else 0, st), 'tgp_kmeans_iteration')
                cur = float(inertia.item())                     # inertia w.r.t. the centroids BEFORE this update
                if it == max_iter or prev - cur <= tol * cur:
                    break
                prev = cur
            if cur < best_inertia:
                best, best_inertia = Cn, cur
    return best.to(cg.dtype)
