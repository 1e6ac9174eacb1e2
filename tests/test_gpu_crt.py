"""Compute mode 'i8crt' (TGP_F64_I8): FP64-accurate contractions on the tcgen05 integer tensor path (csrc/gemm_i8.cuh).

The stand-alone product against an FP64 matmul and, entry for entry, against the CPU restatement oracle/crt_gemm.py; then
the whole ELBO path against the reference fixtures at the FP64 tolerances."""
import warnings

import numpy as np
import pytest
import torch

from oracle import crt_gemm as G
from oracle import tgp_oracle as O
from tests.golden_util import Golden, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _ab(M, N, K, seed, spread=3.0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp(spread * torch.randn(M, 1, generator=g, dtype=torch.float64))
    B = torch.randn(N, K, generator=g, dtype=torch.float64) * torch.exp(spread * torch.randn(N, 1, generator=g, dtype=torch.float64))
    return A, B


@pytest.mark.parametrize('M,N,K', [(128, 256, 128), (100, 37, 53), (455, 100, 100), (257, 513, 1024), (1, 1, 1), (300, 2048, 1024)])
@pytest.mark.parametrize('T', [16, 15])
def test_crt_gemm_matches_fp64_matmul(M, N, K, T):
    from tgp.pytorch_b200.engine import debug_gemm_crt
    A, B = _ab(M, N, K, M + N + K + T)
    out = torch.full((M, N + 3), 7.0, dtype=torch.float64, device=DEV)
    debug_gemm_crt(A.to(DEV), B.to(DEV), out[:, :N], T=T)
    ref = A @ B.t()
    assert rel_err(out[:, :N].cpu(), ref) < 5e-15
    # per-row accuracy too (rows differ in scale by e^3): every row is as good as FP64
    rows = (out[:, :N].cpu() - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)
    assert float(rows.max()) < 1e-13
    assert torch.all(out[:, N:] == 7.0)


def test_crt_gemm_equals_the_cpu_restatement_entry_for_entry():
    from tgp.pytorch_b200.engine import debug_gemm_crt
    A, B = _ab(70, 45, 333, 5)
    A[3] = 0.0
    out = torch.zeros(70, 45, dtype=torch.float64, device=DEV)
    debug_gemm_crt(A.to(DEV), B.to(DEV), out, T=16)
    cpu = G.crt_matmul(A.numpy(), B.numpy(), 16)
    assert np.array_equal(out.cpu().numpy(), cpu)          # same integers, same words, same final rounding


def test_crt_gemm_triangular_lower_and_accumulate():
    from tgp.pytorch_b200.engine import debug_gemm_crt
    g = torch.Generator().manual_seed(9)
    n = 512
    Lo = torch.randn(n, n, generator=g, dtype=torch.float64).tril()
    De = torch.randn(300, n, generator=g, dtype=torch.float64)
    # tri_mode 1: B rows are lower triangular (k <= n): only the needed k-blocks are loaded
    out = torch.zeros(300, n, dtype=torch.float64, device=DEV)
    debug_gemm_crt(De.to(DEV), Lo.to(DEV), out, T=15, tri_mode=1, tri_rows=n)
    assert rel_err(out.cpu(), De @ Lo.t()) < 5e-15
    # tri_mode 2: B[n, k] nonzero only for k >= n
    Up = Lo.t().contiguous()
    out = torch.zeros(300, n, dtype=torch.float64, device=DEV)
    debug_gemm_crt(De.to(DEV), Up.to(DEV), out, T=15, tri_mode=2, tri_rows=n)
    assert rel_err(out.cpu(), De @ Up.t()) < 5e-15
    # lower-only output, accumulated on top of existing values
    S = torch.randn(n, 700, generator=g, dtype=torch.float64)
    out = torch.ones(n, n, dtype=torch.float64, device=DEV)
    debug_gemm_crt(S.to(DEV), S.to(DEV), out, T=16, lower_rows=n, accumulate=True)
    ref = (S @ S.t()).tril() + 1.0
    assert rel_err(out.cpu(), ref) < 5e-15


@pytest.mark.parametrize('mn', [1, 2, 3])
@pytest.mark.parametrize('M,N,K', [(128, 256, 128), (200, 300, 1000), (2048, 1024, 4096), (77, 513, 260)])
def test_crt_gemm_mn_major_operands(M, N, K, mn):
    """Operands handed over transposed (reduction over their ROWS): the tensor core reads the 128-byte-swizzled tile MN-major."""
    from tgp.pytorch_b200.engine import debug_gemm_crt
    A, B = _ab(M, N, K, M + N + K + mn, spread=2.0)
    Ain = A.t().contiguous() if mn & 1 else A
    Bin = B.t().contiguous() if mn & 2 else B
    out = torch.zeros(M, N, dtype=torch.float64, device=DEV)
    debug_gemm_crt(Ain.to(DEV), Bin.to(DEV), out, T=16, mn_major=mn)
    # the transposed operand is scaled per column of what was passed = per row of the logical operand: same accuracy
    assert rel_err(out.cpu(), A @ B.t()) < 5e-15


REG = ['synth_reg_d8_m64_p1', 'boston_tgp_steptanh13_p1', 'boston_svgp_p1', 'power_tgp_sal2_p1', 'boston_tgp_sal2_p0',
       'synth_reg_d8_m1024_p1', 'power_idtgp_nodrop_p1', 'boston_tgp_steptanh102_p1']


@pytest.mark.parametrize('name', REG)
def test_i8crt_mode_reproduces_the_reference_at_fp64_tolerances(name):
    from tests.gpu_util import engine_inputs, make_engine
    from tgp.pytorch_b200 import functional as Fn
    from tests.conftest import record_residuals
    g = Golden(name)
    p = g.oracle_params('train')
    lik, nq = g.meta['likelihood'], g.meta['n_quad']
    eng, theta, rowp, names = make_engine(p, lik, nq, DEV, compute='i8crt')
    ei = engine_inputs(p, DEV)
    X, Y = g.t('X').to(DEV).contiguous(), g.t('Y').view(-1).to(DEV).contiguous()
    leaves = [ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta] + ([rowp] if rowp is not None else [])
    for t in leaves:
        t.requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ELL, KLD, rows, mu, v = Fn.elbo_terms(eng, X, Y, g.meta['N'] / X.shape[0], ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'],
                                              ei['L_raw'], ei['log_var_noise'], theta, rowp)
        (ELL - KLD).backward()
    grads = dict(Z=ei['Z'].grad, raw_lengthscale=ei['raw_ls'].grad, raw_outputscale=ei['raw_os'].grad.view(()), m=ei['m'].grad,
                 L_raw=ei['L_raw'].grad, log_var_noise=ei['log_var_noise'].grad.view(()))
    for i, n in enumerate(names):
        grads[n] = theta.grad[i]
    errs = g.grad_errors(grads)
    vals = dict(ELBO=rel_err((ELL - KLD).detach().cpu(), g.t('ELBO')), mu=rel_err(mu.cpu(), g.t('mu')), v=rel_err(v.cpu(), g.t('v')))
    record_residuals('i8crt:' + name, dict(errs, **vals))
    assert vals['ELBO'] < 1e-10 and vals['mu'] < 1e-10 and vals['v'] < 1e-9, vals
    bad = {k: e for k, e in errs.items() if not e < 1e-10}
    assert not bad, (bad, errs)


@pytest.mark.parametrize('name', ['boston_tgp_steptanh13_p1', 'synth_clf_d16_m48_p1', 'boston_idtgp_drop_p1', 'synth_reg_d8_m1024_p1',
                                  'boston_svgp_jitter'])
def test_class_api_in_i8crt_mode(name):
    """cg.compute = 'i8crt' through sparse_MF_SP.ELBO + backward + test_log_likelihood: same tolerances as the FP64 DMMA mode."""
    from tests.model_util import build_from_golden, set_dropout_mode
    from tests.test_gpu_parity import BERNOULLI_TOL
    from tgp.pytorch_b200.dsp import config as cg
    g = Golden(name)
    old = cg.compute
    try:
        model = build_from_golden(g, DEV)
        cg.compute = 'i8crt'
        set_dropout_mode(model, g)
        X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ELBO, ELL, KLD = model.ELBO(X, Y)
            (-ELBO).backward()
        bern = g.meta['likelihood'] == 'bernoulli'
        vtol = BERNOULLI_TOL['ELBO'] if bern else (1e-7 if 'jitter' in name else 1e-10)
        gtol = BERNOULLI_TOL['grads'] if bern else 1e-10
        assert rel_err(ELBO.detach().cpu(), g.t('ELBO')) < vtol
        assert any(e.compute == 'i8crt' for e in model._engines.values())
        if 'jitter' not in name:
            worst = {}
            for n, prm in model.named_parameters():
                if 'grad:' + n not in g.z.files:
                    worst.update({k: e for k, e in g.grad_errors({'L_raw': -prm.grad.detach()[0]}).items() if k.startswith('L_raw.')})
                    continue
                ref = -g.t('grad:' + n)
                if float(ref.norm()) > 0:
                    worst[n] = rel_err(prm.grad.detach().cpu().reshape(ref.shape), ref)
            bad = {k: e for k, e in worst.items() if not e < gtol}
            assert not bad, bad
        model.set_is_training(False)
        Xt, Yt = g.t('Xte').to(DEV), g.t('Yte').to(DEV)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            lp, _ = model.test_log_likelihood(Xt, Yt.long() if bern else Yt, return_moments=True, Y_std=torch.ones(1, device=DEV) * g.meta['y_std'])
        assert rel_err(lp.double().sum().cpu(), g.t('test_logp')) < (1e-5 if bern else 1e-10)
    finally:
        cg.compute = old
