"""GPU test of the pinned-host minibatch stager: the staged device tensors must equal plain host
indexing for every step, also when compute lags behind the staging (slots are recycled only after their consumer ran)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_stager_matches_host_indexing_under_load():
    from tgp.pytorch_b200.data import PinnedMinibatchStager
    g = torch.Generator().manual_seed(3)
    N, D, rows = 20000, 8, 4096
    X = torch.randn(N, D, generator=g, dtype=torch.float64)
    Y = torch.randn(N, 1, generator=g, dtype=torch.float64)
    perm = torch.randperm(N, generator=g)
    st = PinnedMinibatchStager(X, Y, rows, DEV)
    big = torch.randn(4096, 4096, device=DEV)
    sums = []
    idxs = [perm[(s * 1777) % (N - rows):][:rows] for s in range(12)]
    st.stage(idxs[0])
    for s in range(12):
        xb, yb = st.get()
        if s + 1 < 12:
            st.stage(idxs[s + 1])
        for _ in range(3):                       # keep the compute stream busy so that staging runs ahead of it
            big = (big @ big).tanh_()
        sums.append((xb.sum() + 2.0 * yb.sum(), xb[17].clone(), yb[-1].clone()))
    torch.cuda.synchronize()
    for s, (tot, row17, ylast) in enumerate(sums):
        ref = X[idxs[s]].sum() + 2.0 * Y[idxs[s]].sum()
        assert abs(float(tot) - float(ref)) < 1e-9 * (1.0 + abs(float(ref)))
        assert torch.equal(row17.cpu(), X[idxs[s]][17]) and torch.equal(ylast.cpu(), Y[idxs[s]][-1])


def test_stager_refuses_cpu_device_and_overflow():
    from tgp.pytorch_b200.data import PinnedMinibatchStager
    X, Y = torch.zeros(10, 2, dtype=torch.float64), torch.zeros(10, 1, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        PinnedMinibatchStager(X, Y, 4, 'cpu')
    st = PinnedMinibatchStager(X, Y, 4, DEV)
    with pytest.raises(ValueError):
        st.stage(torch.arange(5))
    st.stage(torch.arange(4)); st.stage(torch.arange(4))
    with pytest.raises(RuntimeError):
        st.stage(torch.arange(4))
    with pytest.raises(RuntimeError):
        PinnedMinibatchStager(X, Y, 4, DEV).get()
