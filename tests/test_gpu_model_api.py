"""GPU tests through the reference-shaped class API (`tgp.pytorch_b200.dsp`): the same calls the reference's trainer
makes (trainers_regression.py:83-92, 317-338) against the fixtures the unmodified reference produced."""
import warnings

import numpy as np
import pytest
import torch

from tests.golden_util import Golden, golden_names, rel_err
from tests.model_util import build_from_golden, set_dropout_mode

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _tols(g):
    bern = g.meta['likelihood'] == 'bernoulli'
    from tests.test_gpu_parity import BERNOULLI_TOL          # measured reason: tests/test_bernoulli_conditioning.py
    return (BERNOULLI_TOL['ELBO'] if bern else 1e-10), (BERNOULLI_TOL['grads'] if bern else 1e-10)


@pytest.mark.parametrize('name', golden_names())
def test_model_elbo_and_named_gradients(name):
    g = Golden(name)
    model = build_from_golden(g, DEV)
    set_dropout_mode(model, g)             # dropout-off fixtures: eval mode; dropout-on fixtures: the recorded masks
    X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
    vtol, gtol = _tols(g)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ELBO, ELL, KLD = model.ELBO(X, Y)
        (-ELBO).backward()                   # what the trainer does: loss = -ELBO
    assert rel_err(ELBO.detach().cpu(), g.t('ELBO')) < vtol
    assert rel_err(ELL.detach().cpu(), g.t('ELL')) < vtol
    assert rel_err(KLD.detach().cpu(), g.t('KLD')) < 1e-12
    if 'jitter' in name:
        return
    worst = {}
    for n, prm in model.named_parameters():
        if 'grad:' + n not in g.z.files:       # BASELINE-size fixtures: the M x M gradient is stored as checksums
            errs = g.grad_errors({'L_raw': -prm.grad.detach()[0]})
            worst.update({k: e for k, e in errs.items() if k.startswith('L_raw.')})
            continue
        ref = -g.t('grad:' + n)
        got = torch.zeros_like(ref) if prm.grad is None else prm.grad.detach().cpu().reshape(ref.shape)
        if float(ref.norm()) == 0.0:
            # exactly-zero reference gradients (symmetric P0 state): ours may carry summation round-off
            assert float(got.norm()) < 1e-12 * abs(float(g.t('ELBO'))), n
            continue
        worst[n] = rel_err(got, ref)
    from tests.conftest import record_residuals
    record_residuals('model_api:' + name, worst)
    if name == 'boston_tgp_steptanh154_p1':      # 60 composed tanh layers: see tests/test_gpu_parity.py
        gtol = 1e-9
    bad = {k: e for k, e in worst.items() if not e < gtol}
    assert not bad, bad


@pytest.mark.parametrize('name', golden_names())
def test_model_marginals_and_standalone_pieces(name):
    g = Golden(name)
    model = build_from_golden(g, DEV)
    set_dropout_mode(model, g)
    X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        mu, v = model.marginal_variational_qf_parameters(X, diagonal=True, is_duvenaud=False)
        assert mu.shape == (1, X.shape[0], 1) and v.shape == (1, X.shape[0], 1)
        assert rel_err(mu.detach().cpu().view(-1), g.t('mu')) < 1e-10
        assert rel_err(v.detach().cpu().view(-1), g.t('v')) < 1e-9
        # ELL from given marginals + stand-alone KLD reproduce the fused ELBO (reference ELBO = ELL.sum() - KLD.sum())
        ell = model.ELL(X, Y, mu.detach(), v.detach())
        kld = model.KLD()
    vtol, _ = _tols(g)
    if 'jitter' not in name:
        assert rel_err((ell.sum() - kld.sum()).detach().cpu(), g.t('ELBO')) < vtol
    assert rel_err(kld.sum().detach().cpu(), g.t('KLD')) < 1e-12


@pytest.mark.parametrize('name', golden_names())
def test_model_test_log_likelihood(name):
    g = Golden(name)
    model = build_from_golden(g, DEV)
    model.set_is_training(False)
    Xt, Yt = g.t('Xte').to(DEV), g.t('Yte').to(DEV)
    lik = g.meta['likelihood']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        lp, mom = model.test_log_likelihood(Xt, Yt.long() if lik == 'bernoulli' else Yt, return_moments=True,
                                            Y_std=torch.ones(1, device=DEV) * g.meta['y_std'], S_MC_NNet=None)
    if lik == 'bernoulli':
        assert rel_err(lp.double().cpu(), g.t('test_logp')) < 1e-5          # scored in float32 by the reference
        assert rel_err(mom[0].double().cpu(), g.t('test_moment0')) < 1e-6
    else:
        assert rel_err(lp.sum().cpu(), g.t('test_logp')) < 1e-10
        ref0 = g.t('test_moment0')
        assert float((mom[0].cpu().view(-1) - ref0).norm()) < 1e-10 * float(ref0.norm()) + 1e-13
        assert rel_err(mom[1].cpu().view(-1), g.t('test_moment1')) < 1e-9


def test_fully_bayesian_mc_dropout_runs_and_is_consistent():
    """MC-dropout test log-lik (sparse_MF_SP.py:764-768): with dropout probability forced to 0 the S_MC mixture of
    identical components must equal the point-estimate value up to the reference's float32 -0.5*log(pi) constant."""
    g = Golden('boston_idtgp_nodrop_p1')
    model = build_from_golden(g, DEV)
    for m in model.modules():
        if 'Dropout' in type(m).__name__:
            m.p = 0.0
    model.set_is_training(False)
    Xt, Yt = g.t('Xte').to(DEV), g.t('Yte').to(DEV)
    ystd = torch.ones(1, device=DEV) * g.meta['y_std']
    lp_pe, _ = model.test_log_likelihood(Xt, Yt, return_moments=False, Y_std=ystd)
    model.be_fully_bayesian(True)
    lp_ba, mom = model.test_log_likelihood(Xt, Yt, return_moments=True, Y_std=ystd, S_MC_NNet=7)
    MB = Xt.shape[0]
    # point estimate subtracts float32(0.5*MB*log pi); the Bayesian branch subtracts MB * float32(0.5*log pi)
    c_pe = float(0.5 * MB * torch.log(torch.tensor(np.pi, dtype=torch.float32)))
    c_ba = MB * float(0.5 * torch.log(torch.tensor(np.pi, dtype=torch.float32)))
    assert abs((float(lp_pe.sum()) + c_pe) - (float(lp_ba.sum()) + c_ba)) < 1e-9 * abs(float(lp_pe.sum()))
    assert rel_err(mom[0].cpu().view(-1), g.t('test_moment0')) < 1e-10


def test_second_forward_before_backward_is_refused():
    g = Golden('boston_svgp_p1')
    model = build_from_golden(g, DEV)
    X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
    E1, _, _ = model.ELBO(X, Y)
    model.ELBO(X, Y)
    with pytest.raises(RuntimeError, match='evaluated again'):
        E1.backward()


def test_predictive_sampling_statistics():
    """sample_from_predictive_distribution (trainers_regression.py:331: S samples per row for coverage): the sample mean /
    variance per row must agree with the quadrature moments of test_log_likelihood within Monte-Carlo error."""
    g = Golden('boston_tgp_sal2_p1')
    model = build_from_golden(g, DEV)
    model.set_is_training(False)
    Xt, Yt = g.t('Xte').to(DEV), g.t('Yte').to(DEV)
    torch.manual_seed(0)
    S = 4000
    samples, f_k, f_0 = model.sample_from_predictive_distribution(Xt, S)
    assert samples.shape == (1, S, Xt.shape[0], 1) and f_k.shape == (1, S * Xt.shape[0])
    _, mom = model.test_log_likelihood(Xt, Yt, return_moments=True, Y_std=torch.ones(1, device=DEV))
    m1, m2 = mom[0].view(-1), mom[1].view(-1)
    sm, sv = samples[0, :, :, 0].mean(0), samples[0, :, :, 0].var(0)
    assert float(((sm - m1).abs() / m2.sqrt()).max()) < 6.0 / S ** 0.5 * 1.5
    assert float((sv / m2 - 1).abs().max()) < 0.25


def test_eval_factorisation_cache_is_transparent():
    """Consecutive no-grad evaluations reuse the factorisation; any parameter update or training call invalidates it."""
    g = Golden('boston_tgp_sal2_p1')
    model = build_from_golden(g, DEV)
    Xt = g.t('Xte').to(DEV)
    with torch.no_grad():
        mu1, v1 = model.marginal_variational_qf_parameters(Xt, diagonal=True, is_duvenaud=False)
        eng = [e for k, e in model._engines.items() if k[0] == 'qf'][0]
        gen = eng.generation
        mu2, v2 = model.marginal_variational_qf_parameters(Xt, diagonal=True, is_duvenaud=False)
        assert eng.prepared_key is not None and torch.equal(mu1, mu2) and torch.equal(v1, v2)
        model.Z.add_(0.01)                                   # in-place update bumps the version -> refactorise
        mu3, _ = model.marginal_variational_qf_parameters(Xt, diagonal=True, is_duvenaud=False)
        assert not torch.equal(mu1, mu3)
        model.Z.sub_(0.01)
    model.set_is_training(True)
    assert eng.prepared_key is None
