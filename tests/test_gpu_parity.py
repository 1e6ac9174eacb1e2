"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle and the golden fixtures.

Tolerances (north_star: 1e-10 in FP64), L2-relative per tensor: 1e-10 on ELBO / ELL / per-row terms / mu / every
gradient, 1e-9 on v (v = s - |a|^2 + |b|^2 cancels; the reference's own cholesky_solve form carries ~cond(K_zz) eps),
1e-12 on KLD.  Exceptions, each with its measured reason: the Bernoulli fixtures (GRAD_TOL_BERNOULLI below and
tests/test_bernoulli_conditioning.py) and the singular-K_zz jitter fixture.  The fixtures include two recorded at
the BASELINE.json sizes (M = 1024 / D = 8 / StepTanhL(1,3) and M = 2048 / D = 16 / Bernoulli / SAL(1)).
"""
import pytest
import torch

from oracle import tgp_oracle as O
from tests.golden_util import Golden, golden_names, rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
# Bernoulli ELL: log(1 - Phi(g)) and asinh(f) = log(f + sqrt(f^2 + 1)) are formed by cancellation in the reference.
# tests/test_bernoulli_conditioning.py measures, against a 50-digit evaluation of the SAME formula, how far the reference's
# own FP64 numbers are from exact: 6.4e-7 (sum) / 1.5e-6 (worst row) on the small fixture and 1.7e-4 / 8e-3 on the
# M = 2048 one.  Two faithful FP64 implementations (host libm vs device erf / log) cannot agree better than a fraction of
# that; the bounds below are ~10x what this path measures against the reference (profiles/r02_parity_residuals.json).
BERNOULLI_TOL = dict(ELBO=5e-6, rows=1e-4, grads=2e-6)      # measured: 6.3e-7 / 8.5e-6 / 1.2e-7 at M = 2048


def _gemm_ref(A, B, al, bl, M, N, K):
    Aop = A[:M, :K] if al == 0 else A[:K, :M].t()
    Bop = B[:N, :K] if bl == 0 else B[:K, :N].t()
    return Aop @ Bop.t()


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (100, 37, 53), (455, 100, 100), (257, 300, 129), (64, 64, 64)])
@pytest.mark.parametrize('al', [0, 1])
@pytest.mark.parametrize('bl', [0, 1])
def test_gemm_layouts(M, N, K, al, bl):
    from tgp.pytorch_b200.engine import debug_gemm
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + al * 2 + bl)
    ld_pad = 3     # odd leading dimensions exercise the scalar (unaligned) load path
    A = torch.randn((M if al == 0 else K), (K if al == 0 else M) + ld_pad, generator=g, dtype=torch.float64).to(DEV)
    B = torch.randn((N if bl == 0 else K), (K if bl == 0 else N) + ld_pad, generator=g, dtype=torch.float64).to(DEV)
    C0 = torch.randn(M, N + 1, generator=g, dtype=torch.float64).to(DEV)
    C = C0.clone()
    debug_gemm(A, B, C, M, N, K, A.stride(0), B.stride(0), C.stride(0), al, bl, alpha=0.5, beta=-2.0)
    ref = 0.5 * _gemm_ref(A, B, al, bl, M, N, K) - 2.0 * C0[:, :N]
    assert rel_err(C[:, :N].cpu(), ref.cpu()) < 1e-13
    assert torch.equal(C[:, N:], C0[:, N:])


def test_gemm_triangular_flags():
    from tgp.pytorch_b200.engine import debug_gemm
    g = torch.Generator().manual_seed(5)
    n = 384
    Lo = torch.randn(n, n, generator=g, dtype=torch.float64).tril().to(DEV)
    Up = torch.randn(n, n, generator=g, dtype=torch.float64).triu().to(DEV)
    De = torch.randn(n, n, generator=g, dtype=torch.float64).to(DEV)
    # a_tri = 1 (lower A), b_tri = 1 (lower B), dense C
    C = torch.zeros(n, n, dtype=torch.float64, device=DEV)
    debug_gemm(Lo, Lo, C, n, n, n, n, n, n, 0, 0, a_tri=1, b_tri=1)
    assert rel_err(C.cpu(), (Lo @ Lo.t()).cpu()) < 1e-13
    # a_tri = 2 via transposed lower, b_tri = 2 via [k][n] lower, lower-only output
    C = torch.full((n, n), 7.0, dtype=torch.float64, device=DEV)
    debug_gemm(Lo, Lo, C, n, n, n, n, n, n, 1, 1, a_tri=2, b_tri=2, c_lower=1)
    ref = (Lo.t() @ Lo).tril() + torch.full((n, n), 7.0, dtype=torch.float64, device=DEV).triu(1)
    assert rel_err(C.cpu(), ref.cpu()) < 1e-13
    # dense x upper ([n][k] upper means k >= n)
    C = torch.zeros(n, n, dtype=torch.float64, device=DEV)
    debug_gemm(De, Up, C, n, n, n, n, n, n, 0, 0, b_tri=2)
    assert rel_err(C.cpu(), (De @ Up.t()).cpu()) < 1e-13


@pytest.mark.parametrize('M,D', [(100, 13), (64, 8), (300, 5), (1024, 8)])
def test_prepare_cholesky_inverse_kl(M, D):
    from tgp.pytorch_b200.engine import Engine, FlowLayout
    g = torch.Generator().manual_seed(M + D)
    p = dict(Z=torch.randn(M, D, generator=g, dtype=torch.float64),
             raw_lengthscale=O.inv_softplus(0.8 + 2 * torch.rand(D, generator=g, dtype=torch.float64)),
             raw_outputscale=O.inv_softplus(torch.tensor(1.5, dtype=torch.float64)),
             m=torch.randn(M, generator=g, dtype=torch.float64),
             L_raw=0.5 * torch.eye(M, dtype=torch.float64) + 0.05 * torch.randn(M, M, generator=g, dtype=torch.float64),
             log_var_noise=torch.tensor(-1.0, dtype=torch.float64), flow=[])
    eng = Engine(M, D, 'gauss_linear', 0, FlowLayout([]), DEV, compute='tf32x3')    # this mode also forms C
    from tests.gpu_util import engine_inputs
    ei = engine_inputs(p, DEV)
    eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'],
                   torch.zeros(0, dtype=torch.float64, device=DEV))
    kl, status = eng.prepare(0.0)
    assert eng.status_reader()() == 0            # host-blocking read of the pivot status (the factorisation may run on its own stream)
    L, Linv, Cm = (t.cpu() for t in eng.export_step())
    assert int(status.item()) == 0               # ... and the device word, in stream order after a consumer has joined
    Kzz = O.rbf_ard(p['Z'], p['Z'], p['raw_lengthscale'], p['raw_outputscale'])
    Lref = torch.linalg.cholesky(Kzz)
    assert rel_err(L, Lref) < 1e-11
    assert rel_err(L @ L.t(), Kzz) < 1e-13
    eye = torch.eye(M, dtype=torch.float64)
    assert float((Linv @ L - eye).abs().max()) < 1e-9
    LS = p['L_raw'].tril()
    assert rel_err(Cm, LS.t() @ torch.linalg.inv(Lref)) < 1e-9
    assert rel_err(kl.cpu(), O.kl_whitened(p)) < 1e-13


def _run_cuda_elbo(g, which='train'):
    """Returns dict with the CUDA path's ELBO terms, marginals, per-row terms and gradients for a fixture."""
    from tests.gpu_util import engine_inputs, make_engine
    from tgp.pytorch_b200 import functional as Fn
    p = g.oracle_params(which)
    lik, nq = g.meta['likelihood'], g.meta['n_quad']
    eng, theta, rowp, names = make_engine(p, lik, nq, DEV)
    ei = engine_inputs(p, DEV)
    X = g.t('X').to(DEV).contiguous()
    Y = g.t('Y').view(-1).to(DEV).contiguous()
    leaves = [ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta]
    if rowp is not None:
        leaves.append(rowp)
    for t in leaves:
        t.requires_grad_(True)
    scale = g.meta['N'] / X.shape[0]
    ELL, KLD, rows, mu, v = Fn.elbo_terms(eng, X, Y, scale, ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'],
                                          None if lik == 'bernoulli' else ei['log_var_noise'], theta, rowp)
    ELBO = ELL - KLD
    ELBO.backward()
    grads = dict(Z=ei['Z'].grad, raw_lengthscale=ei['raw_ls'].grad, raw_outputscale=ei['raw_os'].grad.view(()),
                 m=ei['m'].grad, L_raw=ei['L_raw'].grad)
    if lik != 'bernoulli':
        grads['log_var_noise'] = ei['log_var_noise'].grad.view(())
    for i, n in enumerate(names):
        grads[n] = theta.grad[i]
    return dict(ELBO=ELBO, ELL=ELL, KLD=KLD, rows=rows, mu=mu, v=v, grads=grads, p=p,
                rowp_grad=None if rowp is None else rowp.grad)


@pytest.mark.parametrize('name', golden_names())
def test_elbo_against_reference_fixture(name):
    g = Golden(name)
    import warnings
    from tests.conftest import record_residuals
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = _run_cuda_elbo(g)
    p = g.oracle_params('train')
    lik, nq = g.meta['likelihood'], g.meta['n_quad']
    # per-row expected log-likelihood against the oracle (the reference only exposes the sum)
    rows = O.elbo(g.t('X'), g.t('Y').view(-1), p, g.meta['N'], lik, nq)[3]
    worst = g.grad_errors(out['grads'])
    vals = dict(ELBO=rel_err(out['ELBO'].cpu(), g.t('ELBO')), ELL=rel_err(out['ELL'].cpu(), g.t('ELL')),
                KLD=rel_err(out['KLD'].cpu(), g.t('KLD')), mu=rel_err(out['mu'].cpu(), g.t('mu')),
                v=rel_err(out['v'].cpu(), g.t('v')), rows=rel_err(out['rows'].cpu(), rows))
    record_residuals('fixture:' + name, dict(worst, **vals))
    tol = dict(ELBO=1e-10, ELL=1e-10, KLD=1e-12, mu=1e-10, v=1e-9, rows=1e-10)
    gtol = 1e-10
    if lik == 'bernoulli':
        tol.update(ELBO=BERNOULLI_TOL['ELBO'], ELL=BERNOULLI_TOL['ELBO'], rows=BERNOULLI_TOL['rows'])
        gtol = BERNOULLI_TOL['grads']
    if g.meta.get('expects_jitter'):
        gtol = 1e-7                  # K_zz singular to working precision: the solve amplifies the 1e-8 jitter's round-off
    bad = {k: e for k, e in vals.items() if not e < tol[k]}
    assert not bad, (bad, vals)
    # the deepest shipped flow (15 blocks x 4 tanh steps = 60 composed layers, 270 scalars): single scalars' gradients are
    # sums of cancelling contributions through the composition; measured worst 2.2e-10 on one of the 270 (others <= 4e-11)
    ftol = 1e-9 if name == 'boston_tgp_steptanh154_p1' else gtol
    bad = {k: e for k, e in worst.items() if not e < (ftol if k.startswith('flow') else gtol)}
    assert not bad, (bad, worst)
    if g.meta.get('big'):
        # BASELINE-size fixtures store the M x M gradient as checksums: compare it entry-wise with the oracle (which the
        # CPU suite pins to the same checksums), evaluated here on the host
        og = O.elbo_and_grads(g.t('X'), g.t('Y').view(-1), p, g.meta['N'], lik, nq)[4]
        full = dict(L_raw_full=rel_err(out['grads']['L_raw'].cpu(), og['L_raw']), Z_vs_oracle=rel_err(out['grads']['Z'].cpu(), og['Z']))
        record_residuals('fixture:' + name, full)
        assert all(e < gtol for e in full.values()), full


def test_jitter_ladder_matches_reference():
    import warnings
    g = Golden('boston_svgp_jitter')
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        out = _run_cuda_elbo(g)
    assert any('jitter' in str(x.message) for x in w)
    assert rel_err(out['ELBO'].cpu(), g.t('ELBO')) < 1e-7       # K_zz is singular to working precision here


@pytest.mark.parametrize('name', golden_names())
def test_test_log_likelihood_against_reference_fixture(name):
    from tests.gpu_util import engine_inputs, make_engine
    from tgp.pytorch_b200 import functional as Fn
    g = Golden(name)
    p = g.oracle_params('test')
    lik, nq = g.meta['likelihood'], g.meta['n_quad']
    eng, theta, rowp, _ = make_engine(p, lik, nq, DEV)
    ei = engine_inputs(p, DEV)
    Xt = g.t('Xte').to(DEV).contiguous()
    Yt = g.t('Yte').view(-1).to(DEV).contiguous()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        with torch.no_grad():
            mu, v = Fn.qf_marginals(eng, Xt, ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'])
    eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'],
                   None if lik == 'bernoulli' else ei['log_var_noise'], theta)
    bern_std = v.std().reshape(1) if lik == 'bernoulli' else None
    rp = None if rowp is None else rowp.view(rowp.shape[0], 1, rowp.shape[1]).contiguous()
    logp_rows, m1, m2 = eng.test_rows(mu, v, Yt, rp, 1, g.meta['y_std'], bern_std)
    MB = Xt.shape[0]
    if lik == 'gauss_nonlinear':
        logp = logp_rows.sum().cpu() - 0.5 * MB * torch.log(O.PI_F32)
        assert rel_err(logp, g.t('test_logp')) < 1e-10
        # P0 fixtures have m = 0, so the first moment is pure round-off (~1e-17): absolute floor next to the relative bound
        assert float((m1.cpu() - g.t('test_moment0')).norm()) < 1e-10 * float(g.t('test_moment0').norm()) + 1e-13
        assert rel_err(m2.cpu(), g.t('test_moment1')) < 1e-9
    elif lik == 'gauss_linear':
        assert rel_err(logp_rows.sum().cpu(), g.t('test_logp')) < 1e-10
        assert float((m1.cpu() - g.t('test_moment0')).norm()) < 1e-10 * float(g.t('test_moment0').norm()) + 1e-13
        assert rel_err(m2.cpu(), g.t('test_moment1')) < 1e-9
    else:
        ref = g.t('test_moment0')                      # (MB, 2) probabilities, computed in FP32 by the reference
        assert rel_err(m1.cpu(), ref[:, 1]) < 1e-6


def test_row_shards_sum_to_single_gpu():
    """SURVEY.md §8e on one device: two row slices with the global scale reproduce the single-call ELBO and grads."""
    from tests.gpu_util import engine_inputs, make_engine
    g = Golden('synth_reg_d8_m64_p1')
    full = _run_cuda_elbo(g)
    p = g.oracle_params('train')
    eng, theta, rowp, names = make_engine(p, g.meta['likelihood'], g.meta['n_quad'], DEV)
    ei = engine_inputs(p, DEV)
    X = g.t('X').to(DEV).contiguous()
    Y = g.t('Y').view(-1).to(DEV).contiguous()
    scale = g.meta['N'] / X.shape[0]
    eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    kl, _ = eng.prepare(0.0)
    total = eng.new_reduce_buffer()
    half = X.shape[0] // 2 + 3
    for sl in (slice(0, half), slice(half, X.shape[0])):
        rb = eng.new_reduce_buffer()
        Xs, Ys = X[sl].contiguous(), Y[sl].contiguous()
        mu, v = eng.qf_forward(Xs)
        rows, g_mu, g_v, _ = eng.ell_forward(mu, v, Ys, None, scale, rb)
        eng.qf_backward(Xs, g_mu, g_v, rb)
        total += rb                                     # what the NCCL all-reduce does across ranks
    out = eng.chain_backward(total, 1.0, -1.0)
    ELBO = scale * total[0] - kl[0]
    assert rel_err(ELBO.cpu(), full['ELBO'].detach().cpu()) < 1e-13
    assert rel_err(out['Z'].cpu(), full['grads']['Z'].cpu()) < 1e-11
    assert rel_err(out['L_raw'].cpu(), full['grads']['L_raw'].cpu()) < 1e-11
    assert rel_err(out['raw_ls'].cpu(), full['grads']['raw_lengthscale'].cpu()) < 1e-11
