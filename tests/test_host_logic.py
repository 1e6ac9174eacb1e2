"""CPU tests of the host side: the dsp mirror builds reference-shaped modules (same parameter names), flow
descriptors pack as the kernels expect, the C-ABI library loads and exports every declared symbol, and the product
refuses to run without a GPU instead of falling back."""
import os
import re

import pytest
import torch

from tests.golden_util import Golden, golden_names
from tests.model_util import build_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from tgp.pytorch_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'tgp_b200.h')).read()
    declared = set(re.findall(r'\b(tgp_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    for name in declared:
        assert hasattr(lib, name), 'library does not export %s' % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert lib.tgp_version() >= 100


def test_cabi_rejects_bad_descriptions_without_a_gpu():
    from tgp.pytorch_b200 import _lib
    lib = _lib.load()
    m = _lib.TgpModel()
    m.M, m.D, m.likelihood, m.n_quad = 0, 3, 1, 100
    assert lib.tgp_step_workspace_bytes(m) == 0
    assert b'positive' in lib.tgp_last_error()
    m.M = 16
    m.dtype = 7
    assert lib.tgp_step_workspace_bytes(m) == 0
    m.dtype = _lib.TGP_F32
    assert lib.tgp_step_workspace_bytes(m) > 0
    m.dtype = _lib.TGP_F64
    assert lib.tgp_step_workspace_bytes(m) > 0
    lay = _lib.TgpReduceLayout()
    assert lib.tgp_reduce_layout(m, lay) == 0
    assert lay.total == lay.Cbar + 64 * 64 and lay.Gbar % 2 == 0


@pytest.mark.parametrize('name', golden_names())
def test_models_reproduce_reference_parameter_names(name):
    g = Golden(name)
    model = build_from_golden(g, 'cpu')          # asserts the name sets are identical
    assert model.M == g.meta['M'] and model.out_dim == 1


def test_flow_descriptor_packing():
    from tgp.pytorch_b200.dsp import config as cg
    cg.set_maximum_precission()
    from tgp.pytorch_b200.dsp.flows import StepTanhL, SAL
    from tgp.pytorch_b200.dsp.models.flow import instance_flow
    from tgp.pytorch_b200.engine import FlowLayout
    fl = instance_flow(StepTanhL(2, 3, add_f0=True))
    layers, glob, rows = fl.describe()
    lay = FlowLayout(layers)
    assert [l['kind'] for l in lay.layers] == ['tanh_step', 'affine', 'tanh_step', 'affine']
    assert lay.n_theta == len(glob) == 2 * (3 * 4 + 2) and lay.n_rowparams == 0 and not rows
    assert [l['p0'] for l in lay.layers] == [0, 12, 14, 26]
    idf = instance_flow(SAL(2, input_dependent=True, input_dim=4, hidden_dim=5, dropout=0.25, num_hidden_layers=2))
    idf.turn_off_initializer_parameters()
    layers, glob, rows = idf.describe(torch.randn(7, 4, dtype=torch.float64))
    lay = FlowLayout(layers)
    assert lay.n_rowparams == 4 and lay.n_theta == 4 and len(rows) == 4 and rows[0].shape == (7,)
    assert [(l['per_row'], l['p0']) for l in lay.layers] == [(True, 0), (False, 0), (True, 2), (False, 2)]


def test_flow_modules_match_oracle_forward():
    """The torch `forward` of the flow modules (used by initialisers / sampling) agrees with the oracle's flows."""
    from oracle import tgp_oracle as O
    from tgp.pytorch_b200.dsp import config as cg
    cg.set_maximum_precission()
    from tgp.pytorch_b200.dsp.flows import StepTanhL, SAL
    from tgp.pytorch_b200.dsp.models.flow import instance_flow
    f = torch.linspace(-3, 3, 41, dtype=torch.float64)
    for spec in (StepTanhL(2, 3, add_f0=True), SAL(2, init_random=True)):
        fl = instance_flow(spec)
        layers = []
        for sub in fl.flow_arr:
            kind = type(sub).__name__
            if kind == 'AffineFlow':
                layers.append(('affine', sub.a, sub.b, sub.set_restrictions))
            elif kind == 'StepFlow':
                layers.append(('tanh_step', [(s.a, s.b, s.c, s.d) for s in sub.flow_arr], sub.add_init_f0))
            else:
                layers.append(('sal', sub.a, sub.b, sub.set_restrictions, sub.add_init_f0))
        assert torch.allclose(fl(f), O.apply_flow(layers, f), rtol=1e-14, atol=1e-14)


def test_no_cpu_fallback():
    g = Golden('boston_svgp_p1')
    model = build_from_golden(g, 'cpu')
    with pytest.raises(RuntimeError, match='no CPU path'):
        model.ELBO(g.t('X'), g.t('Y'))


def test_minibatch_stager_has_no_cpu_path():
    """The pinned-host stager copies into CUDA memory; asking it for a CPU device must fail loudly, not degrade."""
    import pytest
    from tgp.pytorch_b200.data import PinnedMinibatchStager
    X, Y = torch.zeros(8, 2, dtype=torch.float64), torch.zeros(8, 1, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        PinnedMinibatchStager(X, Y, 4, 'cpu')


def test_library_options_and_workspace_sizing_without_a_gpu():
    """Option ids in the ctypes binding equal the header's enum; the row-chunk option is validated and changes the batch
    workspace size it documents (pure host code: no kernel is launched)."""
    from tgp.pytorch_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'tgp_b200.h')).read()
    enum = dict(re.findall(r'(TGP_OPT_[A-Z_]+)\s*=\s*(\d+)', header))
    assert int(enum['TGP_OPT_FUSED_FORWARD']) == _lib.OPT_FUSED_FORWARD
    assert int(enum['TGP_OPT_ROW_CHUNK']) == _lib.OPT_ROW_CHUNK
    m = _lib.TgpModel()
    m.dtype, m.M, m.D, m.likelihood, m.n_quad = _lib.TGP_F64, 256, 4, 1, 30
    R = 100000
    try:
        assert lib.tgp_set_option(_lib.OPT_ROW_CHUNK, 8192) == 0
        small = lib.tgp_batch_workspace_bytes(m, R)
        assert lib.tgp_set_option(_lib.OPT_ROW_CHUNK, 65536) == 0
        large = lib.tgp_batch_workspace_bytes(m, R)
        # [A|B] (2M) and K_xz (M) are per row; the two staging buffers (M each) are per chunk
        assert small >= 8 * (3 * 256 * R + 2 * 256 * 8192) and large - small == 8 * 2 * 256 * (65536 - 8192)
        assert lib.tgp_set_option(_lib.OPT_ROW_CHUNK, 64) != 0 and b'row chunk' in lib.tgp_last_error()
        assert lib.tgp_set_option(99, 1) != 0
    finally:
        assert lib.tgp_set_option(_lib.OPT_ROW_CHUNK, 32768) == 0


def test_split_csv_reader_standardises_like_the_reference(tmp_path):
    """data.load_split_csv: the airline-style on-disk format (header-less CSV + splits_idx pickle,
    regression_datasets.py:95-192) with the training-statistics standardisation of data.py:260-299."""
    import pickle
    import numpy as np
    from tgp.pytorch_b200.data import load_split_csv
    rng = np.random.default_rng(0)
    data = rng.normal(size=(200, 6)) * np.array([1, 10, 0.1, 5, 2, 3]) + np.array([0, 5, -1, 2, 0, 7])
    csv = tmp_path / 'airline.csv'
    np.savetxt(csv, data, delimiter=',')
    perm = rng.permutation(200)
    with open(tmp_path / 'splits_idx_airline.pkl', 'wb') as fh:
        pickle.dump({'seed_3': {'train': perm[:150], 'test': perm[150:]}}, fh)
    X_tr, Y_tr, X_te, Y_te, Y_std = load_split_csv(str(csv), str(tmp_path / 'splits_idx_airline.pkl'), 3)
    raw = np.loadtxt(csv, delimiter=',')
    tr, te = raw[perm[:150]], raw[perm[150:]]
    xm, xs = tr[:, :-1].mean(0), tr[:, :-1].std(0) + 1e-15
    ym, ys = tr[:, -1].mean(), tr[:, -1].std() + 1e-15
    assert np.allclose(X_tr.numpy(), (tr[:, :-1] - xm) / xs, rtol=0, atol=1e-14)
    assert np.allclose(X_te.numpy(), (te[:, :-1] - xm) / xs, rtol=0, atol=1e-14)
    assert np.allclose(Y_te.numpy()[:, 0], (te[:, -1] - ym) / ys, rtol=0, atol=1e-14)
    assert abs(float(Y_std) - ys) < 1e-15 and Y_tr.shape == (150, 1)
    with pytest.raises(ValueError, match='md5'):
        load_split_csv(str(csv), str(tmp_path / 'splits_idx_airline.pkl'), 3, md5sum='0' * 32)
