"""Evaluation bundle: posterior-predictive samples and interval coverage from the marginals of the
test-NLL pass (tgp_coverage_rows) — quantiles against numpy.quantile on the kernel's own samples (exact), the samples'
distribution against the analytic predictive moments, and the class-level bundle against the reference-shaped slow path."""
import numpy as np
import pytest
import torch

from tests.golden_util import Golden, rel_err
from tests.model_util import build_from_golden

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('name,S', [('boston_tgp_steptanh13_p1', 100), ('boston_svgp_p1', 128), ('power_tgp_sal2_p1', 37)])
def test_quantiles_and_coverage_match_numpy_on_the_same_samples(name, S):
    g = Golden(name)
    model = build_from_golden(g, DEV)
    model.set_is_training(False)
    Xt, Yt = g.t('Xte').to(DEV), g.t('Yte').to(DEV)
    out = model.evaluation_bundle(Xt, Yt, torch.ones(1, device=DEV) * g.meta['y_std'], S=S, want_samples=True)
    smp = out['samples'][0].cpu().numpy()                     # (MB, S)
    assert smp.shape == (Xt.shape[0], S) and np.isfinite(smp).all()
    q = np.quantile(smp, [0.025, 0.975], axis=1)              # what the reference computes on the host
    assert np.allclose(out['q_lo'][0].cpu().numpy(), q[0], rtol=1e-13, atol=1e-13)
    assert np.allclose(out['q_hi'][0].cpu().numpy(), q[1], rtol=1e-13, atol=1e-13)
    y = Yt[:, 0].cpu().numpy()
    assert int(out['coverage'][0].item()) == int(np.logical_and(y >= q[0], y <= q[1]).sum())
    # the first two outputs are test_log_likelihood's
    assert rel_err(out['log_p_y'].sum().cpu(), g.t('test_logp')) < 1e-10
    assert rel_err(out['m1'].cpu().view(-1), g.t('test_moment0')) < 1e-9


def test_samples_follow_the_predictive_distribution():
    """Many samples per row: their mean / variance agree with the quadrature moments (m1, m2) within Monte-Carlo error, and a
    second call continues the Philox stream (different samples)."""
    g = Golden('boston_tgp_sal2_p1')
    model = build_from_golden(g, DEV)
    model.set_is_training(False)
    Xt, Yt = g.t('Xte').to(DEV), g.t('Yte').to(DEV)
    ystd = torch.ones(1, device=DEV)
    runs = [model.evaluation_bundle(Xt, Yt, ystd, S=128, want_samples=True) for _ in range(40)]
    smp = torch.cat([r['samples'][0] for r in runs], dim=1)            # (MB, 5120)
    assert not torch.equal(runs[0]['samples'], runs[1]['samples'])
    m1, m2 = runs[0]['m1'].view(-1), runs[0]['m2'].view(-1)
    n = smp.shape[1]
    z = (smp.mean(1) - m1) / (m2 / n).sqrt()
    assert float(z.abs().max()) < 5.0, float(z.abs().max())
    assert float((smp.var(1) / m2 - 1).abs().max()) < 0.25
    # nominal coverage: about 95 % of the rows' OWN predictive samples fall inside the interval
    inside = ((smp >= runs[0]['q_lo'][0][:, None]) & (smp <= runs[0]['q_hi'][0][:, None])).double().mean()
    assert 0.92 < float(inside) < 0.97
