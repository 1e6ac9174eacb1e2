"""Device k-means for the inducing-point initialisation (tgp_kmeans_iteration / dsp.utils.KMEANS_device)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _blobs(n, d, k, seed):
    g = torch.Generator().manual_seed(seed)
    centres = 4.0 * torch.randn(k, d, generator=g, dtype=torch.float64)
    idx = torch.randint(0, k, (n,), generator=g)
    return centres[idx] + 0.3 * torch.randn(n, d, generator=g, dtype=torch.float64)


@pytest.mark.parametrize('n,d,k', [(5000, 8, 64), (3000, 13, 100), (20000, 4, 300), (1000, 30, 17)])
def test_fixed_point_and_inertia_against_sklearn(n, d, k):
    from sklearn.cluster import KMeans
    from tgp.pytorch_b200.dsp import config as cg
    from tgp.pytorch_b200.dsp.utils import KMEANS_device
    cg.set_maximum_precission()
    X = _blobs(n, d, k, n + d + k)
    Z = KMEANS_device(X.to(DEV), k, n_init=3, seed=1, max_iter=500, tol=0.0).cpu()      # run to the Lloyd fixed point
    assert Z.shape == (k, d) and torch.isfinite(Z).all()
    # Lloyd fixed point: every centroid is the mean of the points nearest to it (empty clusters keep their centre)
    dist = torch.cdist(X, Z)
    a = dist.argmin(1)
    inertia = float((dist.min(1).values ** 2).sum())
    for c in range(k):
        pts = X[a == c]
        if len(pts):
            assert float((pts.mean(0) - Z[c]).abs().max()) < 1e-9 * (1.0 + float(Z[c].abs().max()))
    ref = KMeans(n_clusters=k, init='k-means++', n_init=3, random_state=0).fit(X.numpy())
    assert inertia < 1.35 * ref.inertia_, (inertia, ref.inertia_)


def test_one_iteration_matches_a_numpy_lloyd_step():
    import ctypes as C
    from tgp.pytorch_b200 import _lib
    lib = _lib.load()
    X = _blobs(4097, 8, 10, 3)
    C0 = X[:33].clone()
    Xd, Cd = X.to(DEV), C0.to(DEV).clone()
    assign = torch.empty(X.shape[0], dtype=torch.int32, device=DEV)
    sums = torch.empty(33, 8, dtype=torch.float64, device=DEV)
    counts = torch.empty(33, dtype=torch.float64, device=DEV)
    inertia = torch.empty(1, dtype=torch.float64, device=DEV)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.tgp_kmeans_iteration(Xd.data_ptr(), X.shape[0], 8, Cd.data_ptr(), 33, assign.data_ptr(), sums.data_ptr(),
                                        counts.data_ptr(), inertia.data_ptr(), 1, st), 'tgp_kmeans_iteration')
    d2 = ((X[:, None, :] - C0[None, :, :]) ** 2).sum(-1)
    a = d2.argmin(1)
    assert torch.equal(assign.cpu().long(), a)
    assert abs(float(inertia.item()) - float(d2.min(1).values.sum())) < 1e-9 * float(d2.min(1).values.sum())
    new = torch.stack([X[a == c].mean(0) if (a == c).any() else C0[c] for c in range(33)])
    assert float((Cd.cpu() - new).abs().max()) < 1e-12
