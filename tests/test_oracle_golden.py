"""O2 (oracle/tgp_oracle.py) must reproduce every fixture the unmodified reference produced (O1)."""
import pytest
import torch

from oracle import tgp_oracle as O
from tests import golden_util
from tests.golden_util import Golden, golden_names, rel_err

TOL = 1e-11      # L2-relative; the oracle replays the reference's own torch ops, so it should be ~1e-14


@pytest.mark.parametrize('name', golden_names())
def test_elbo_marginals_and_grads(name):
    g = Golden(name)
    p = g.oracle_params('train')
    X, Y = g.t('X'), g.t('Y').view(-1)
    lik, nq = g.meta['likelihood'], g.meta['n_quad']
    mu, v = O.qf_marginals(X, p)
    assert rel_err(mu, g.t('mu')) < TOL
    assert rel_err(v, g.t('v')) < TOL
    E, ELL, KLD, rows, grads = O.elbo_and_grads(X, Y, p, g.meta['N'], lik, nq)
    assert rel_err(E, g.t('ELBO')) < TOL
    assert rel_err(ELL, g.t('ELL')) < TOL
    assert rel_err(KLD, g.t('KLD')) < TOL
    errs = g.grad_errors(grads)
    bad = {k: e for k, e in errs.items() if not e < 1e-9}
    assert not bad, (bad, errs)


@pytest.mark.parametrize('name', golden_names())
def test_test_log_likelihood_and_moments(name):
    g = Golden(name)
    p = g.oracle_params('test')
    Xt, Yt = g.t('Xte'), g.t('Yte').view(-1)
    lik = g.meta['likelihood']
    lp, mom = O.test_log_likelihood(Xt, Yt, p, torch.tensor(g.meta['y_std'], dtype=torch.float64), lik,
                                    g.meta['n_quad'])
    assert rel_err(lp, g.t('test_logp')) < TOL
    if lik == 'bernoulli':
        assert rel_err(mom[0].double(), g.t('test_moment0')) < 1e-6      # reference computes these in FP32
    else:
        assert rel_err(mom[0], g.t('test_moment0')) < TOL
        assert rel_err(mom[1], g.t('test_moment1')) < TOL


def test_identity_flow_kat():
    """SAL at its initial parameters is the identity: TGP quadrature ELBO == SVGP closed form."""
    a, b = Golden('boston_svgp_p0'), Golden('boston_tgp_sal2_p0')
    assert abs(float(a.z['ELBO']) - float(b.z['ELBO'])) < 1e-9 * abs(float(a.z['ELBO']))
    assert abs(float(a.z['ELBO']) - (-6294.0125254488)) < 1e-6     # value recorded in BASELINE.md §4


def test_jitter_fixture_takes_the_ladder():
    g = Golden('boston_svgp_jitter')
    p = g.oracle_params('train')
    Kzz = O.rbf_ard(p['Z'], p['Z'], p['raw_lengthscale'], p['raw_outputscale'])
    _, _, jit = O.psd_safe_cholesky(Kzz)
    assert jit > 0


def test_row_shards_sum_to_single():
    g = Golden('synth_reg_d8_m64_p1')
    p = g.oracle_params('train')
    X, Y = g.t('X'), g.t('Y').view(-1)
    E1 = O.elbo(X, Y, p, g.meta['N'], g.meta['likelihood'], g.meta['n_quad'])[0]
    for world in (2, 4, 8):
        Ew = O.elbo_rows_sharded(X, Y, p, g.meta['N'], world, g.meta['likelihood'], g.meta['n_quad'])[0]
        assert rel_err(Ew, E1) < 1e-12


def _mlp_weights(g, li, net):
    """[(W, b), ...] of NNets_<net> of input-dependent layer `li`, by the reference's parameter names."""
    pre = 'param:G_matrix.0.flow_arr.%d.NNets_%s.' % (li, net)
    ks = sorted({int(k[len(pre):].split('.')[0]) for k in g.z.files if k.startswith(pre)})
    return [(g.t(pre + '%d.forward_lin.0.weight' % k), g.t(pre + '%d.forward_lin.0.bias' % k)) for k in ks]


@pytest.mark.parametrize('name,act,p', [('boston_idtgp_drop_p1', 'tanh', 0.5), ('power_idtgp_drop_p1', 'relu', 0.25)])
def test_dropout_fixture_masks_reproduce_the_per_row_flow_parameters(name, act, p):
    """ID_TGP recorded in TRAINING mode (MC-dropout active): the oracle's MLP with the recorded keep-masks must give the
    per-row a(x), b(x) the reference evaluated (stored as the fixture's per-row flow layers)."""
    g = Golden(name)
    assert g.meta['dropout']
    X = g.t('X')
    for li, lay in enumerate(g.meta['flow_train']):
        if lay[0] != 'sal' or 'dropmask:%d' % li not in g.z.files:
            continue
        masks = g.t('dropmask:%d' % li, torch.uint8)                    # (2, L, rows, H)
        for ni, net in enumerate('ab'):
            out = O.flow_mlp(X, _mlp_weights(g, li, net), act, p, [masks[ni, l] for l in range(masks.shape[1])])
            assert rel_err(out, g.t(lay[1 + ni])) < 1e-13, (li, net)


@pytest.mark.parametrize('name', golden_util.multiclass_names())
def test_multiclass_monte_carlo_elbo_grads_and_probabilities(name):
    """Softmax likelihood, one GP per class, with the reference's recorded N(0,1) draws
    (likelihoods/MulticlassCategorical.py:51-151): ELBO / ELL / KLD, marginals, every gradient, class probabilities, test NLL."""
    from oracle import tgp_oracle as O
    g = golden_util.MulticlassGolden(name)
    plist = g.plist()
    X, y = g.t('X'), g.t('Y').reshape(-1)
    leaves = [O.leaf_params(p) for p in plist]
    for lv in leaves:
        lv.pop('log_var_noise')
        for t in lv.values():
            t.requires_grad_(True)
    E, ELL, KLD, rows, mu, v = O.elbo_multiclass(X, y, plist, g.meta['N'], g.t('eps'))
    assert golden_util.rel_err(E, g.t('ELBO')) < 1e-12
    assert golden_util.rel_err(ELL, g.t('ELL')) < 1e-12
    assert golden_util.rel_err(KLD, g.t('KLD')) < 1e-12
    assert golden_util.rel_err(mu.detach(), g.t('mu')) < 1e-11
    assert golden_util.rel_err(v.detach(), g.t('v')) < 1e-10
    E.backward()
    for c, lv in enumerate(leaves):
        ref = g.ref_grads_of_class(c)
        assert set(ref) == set(lv)
        for k, t in lv.items():
            assert golden_util.rel_err(t.grad, ref[k]) < 1e-9, (c, k)
    with torch.no_grad():
        Xte = g.t('Xte')
        mv = [O.qf_marginals(Xte, p) for p in plist]
        P = O.probs_mc_softmax(torch.stack([a for a, _ in mv]), torch.stack([b for _, b in mv]), [p['flow'] for p in plist],
                               g.t('eps_te'))
        assert golden_util.rel_err(P, g.t('test_probs')) < 1e-11
        nll = -torch.log(P.float().gather(1, g.t('Yte').long().view(-1, 1))).mean()          # scored in float32 (sparse_MF_SP.py:813)
        assert abs(float(-(nll * Xte.shape[0])) - float(g.t('test_logp'))) < 1e-5 * abs(float(g.t('test_logp')))
