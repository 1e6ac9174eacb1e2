"""The single-call C-ABI (tgp_create / tgp_bind_workspace / tgp_elbo_fwd / tgp_elbo_bwd / tgp_test_nll_fwd) against the
reference fixtures, and its all-reduce callback against a two-shard evaluation on one device."""
import pytest
import torch

from oracle import tgp_oracle as O
from tests.golden_util import Golden, rel_err
from tests.gpu_util import engine_inputs, make_engine

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _session(g, max_rows):
    from tgp.pytorch_b200.session import ElboSession
    p = g.oracle_params('train')
    eng, theta, rowp, names = make_engine(p, g.meta['likelihood'], g.meta['n_quad'], DEV)
    assert rowp is None
    return ElboSession(eng, max_rows), engine_inputs(p, DEV), theta, names


@pytest.mark.parametrize('name', ['synth_reg_d8_m64_p1', 'boston_tgp_steptanh13_p1', 'boston_svgp_p1', 'synth_reg_d8_m1024_p1'])
def test_single_call_forward_backward_match_the_reference(name):
    g = Golden(name)
    X, Y = g.t('X').to(DEV).contiguous(), g.t('Y').view(-1).to(DEV).contiguous()
    ses, ei, theta, names = _session(g, X.shape[0] + 5)
    args = (ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    out = ses.forward(X, Y, g.meta['N'] / X.shape[0], *args)
    assert int(out['status'].item()) == 0
    assert rel_err((out['terms'][0] - out['terms'][1]).cpu(), g.t('ELBO')) < 1e-10
    assert rel_err(out['terms'][1].cpu(), g.t('KLD')) < 1e-12
    assert rel_err(out['mu'].cpu(), g.t('mu')) < 1e-10
    grads = ses.backward(torch.tensor([1.0, -1.0], dtype=torch.float64, device=DEV))          # d(ELL - KL)
    got = dict(Z=grads['Z'], raw_lengthscale=grads['raw_ls'], raw_outputscale=grads['raw_os'].view(()), m=grads['m'],
               L_raw=grads['L_raw'], log_var_noise=grads['log_var_noise'].view(()))
    for i, n in enumerate(names):
        got[n] = grads['theta'][i]
    errs = g.grad_errors(got)
    assert all(e < 1e-10 for e in errs.values()), errs
    ses.close()


def test_backward_requires_the_matching_forward_and_a_bound_workspace():
    from tgp.pytorch_b200 import _lib
    g = Golden('synth_reg_d8_m64_p1')
    X, Y = g.t('X').to(DEV).contiguous(), g.t('Y').view(-1).to(DEV).contiguous()
    ses, ei, theta, _ = _session(g, 256)
    args = (ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    with pytest.raises(ValueError, match='max_rows'):
        ses.forward(X, Y, 1.0, *args)                      # 512 rows > 256
    ses.forward(X[:200].contiguous(), Y[:200].contiguous(), 1.0, *args)
    ses._batch.R = 100
    with pytest.raises(ValueError, match='same batch'):
        ses.backward(torch.ones(2, dtype=torch.float64, device=DEV))
    small = torch.empty(1024, dtype=torch.uint8, device=DEV)
    lib = _lib.load()
    assert lib.tgp_bind_workspace(ses.handle, small.data_ptr(), 1024) != 0
    assert b'smaller' in lib.tgp_last_error()


def test_allreduce_callback_sums_row_shards():
    """Two sessions hold the two row slices of a minibatch; the callback of the second adds the first one's packed buffer
    (what ncclAllReduce does across ranks): gradients equal the single-call evaluation of the whole batch."""
    g = Golden('synth_reg_d8_m64_p1')
    X, Y = g.t('X').to(DEV).contiguous(), g.t('Y').view(-1).to(DEV).contiguous()
    scale = g.meta['N'] / X.shape[0]
    one, ei, theta, _ = _session(g, 512)
    args = (ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    gd = torch.tensor([1.0, -1.0], dtype=torch.float64, device=DEV)
    one.forward(X, Y, scale, *args)
    ref = one.backward(gd)
    a, _, _, _ = _session(g, 512)
    b, _, _, _ = _session(g, 512)
    a.forward(X[:300].contiguous(), Y[:300].contiguous(), scale, *args)
    b.forward(X[300:].contiguous(), Y[300:].contiguous(), scale, *args)
    stash = {}
    a.backward(gd, allreduce=lambda t: stash.__setitem__('a', t.clone()))       # rank 0's contribution
    got = b.backward(gd, allreduce=lambda t: t.add_(stash['a']))                 # rank 1 receives the sum
    for k in ('Z', 'raw_ls', 'raw_os', 'm', 'L_raw', 'log_var_noise', 'theta'):
        assert rel_err(got[k].cpu(), ref[k].cpu()) < 1e-12, k


def test_single_call_test_nll():
    g = Golden('boston_tgp_steptanh13_p1')
    Xt, Yt = g.t('Xte').to(DEV).contiguous(), g.t('Yte').view(-1).to(DEV).contiguous()
    from tgp.pytorch_b200.session import ElboSession
    p = g.oracle_params('test')
    eng, theta, _, _ = make_engine(p, 'gauss_nonlinear', g.meta['n_quad'], DEV)
    ei = engine_inputs(p, DEV)
    ses = ElboSession(eng, 64)
    args = (ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    logp, m1, m2, mu, v, status = ses.test_nll(Xt, Yt, *args, y_std=g.meta['y_std'])
    assert int(status.item()) == 0
    lp = logp.sum().cpu() - 0.5 * Xt.shape[0] * torch.log(O.PI_F32)
    assert rel_err(lp, g.t('test_logp')) < 1e-10
    assert rel_err(m2.cpu(), g.t('test_moment1')) < 1e-9
    logp2 = ses.test_nll(Xt, Yt, *args, y_std=g.meta['y_std'], refactor=False)[0]      # factorisation reused
    assert torch.equal(logp, logp2)
