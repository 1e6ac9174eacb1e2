"""GPU parity of the Monte-Carlo softmax likelihood: the kernel through the C-ABI against the oracle,
and `sparse_MF_SP` with `MulticlassCategorical` through the class API against fixtures of the unmodified reference
(likelihoods/MulticlassCategorical.py:51-151; one GP per class, sparse_MF_SP.py:552-626)."""
import ctypes as C
import warnings

import numpy as np
import pytest
import torch

from oracle import tgp_oracle as O
from tests.conftest import record_residuals
from tests.golden_util import MulticlassGolden, multiclass_names, rel_err
from tests.gpu_util import flow_layout_and_params

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-10


def _kernel(g, mu, v, y, eps, want_probs=False):
    from tgp.pytorch_b200 import _lib
    from tgp.pytorch_b200.dsp.likelihoods._rows import _mc_model
    lib = _lib.load()
    plist = g.plist()
    packs = [flow_layout_and_params(p['flow'], DEV) for p in plist]
    layout, names = packs[0][0], packs[0][3]
    theta = torch.stack([p[1] for p in packs]).contiguous()
    Cn, R = mu.shape
    S = eps.shape[0]
    d = dict(dtype=torch.float64, device=DEV)
    rows, g_mu, g_v = torch.empty(R, **d), torch.empty(Cn, R, **d), torch.empty(Cn, R, **d)
    dth, probs = torch.zeros_like(theta), torch.empty(R, Cn, **d)
    dev = lambda t: t.to(DEV).double().contiguous()        # noqa: E731
    mu_d, v_d, y_d, e_d = dev(mu), dev(v), dev(y), dev(eps)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.tgp_mc_softmax_rows(_mc_model(layout), Cn, S, R, mu_d.data_ptr(), v_d.data_ptr(), y_d.data_ptr(), e_d.data_ptr(),
                                       theta.data_ptr(), 1, rows.data_ptr(), g_mu.data_ptr(), g_v.data_ptr(), dth.data_ptr(),
                                       probs.data_ptr() if want_probs else None, st), 'tgp_mc_softmax_rows')
    torch.cuda.synchronize()
    return rows.cpu(), g_mu.cpu(), g_v.cpu(), dth.cpu(), probs.cpu(), names


@pytest.mark.parametrize('name', multiclass_names())
def test_kernel_rows_gradients_and_probabilities_against_the_oracle(name):
    g = MulticlassGolden(name)
    plist = g.plist()
    mu, v = g.t('mu').clone().requires_grad_(True), g.t('v').clone().requires_grad_(True)
    y, eps = g.t('Y').reshape(-1), g.t('eps')
    flows = [p['flow'] for p in plist]
    leaves = [dict(sum((O.layer_leaves(lay, 'flow%d' % i) for i, lay in enumerate(f)), [])) for f in flows]
    for lv in leaves:
        for t in lv.values():
            t.requires_grad_(True)
    rows_o = O.ell_rows_mc_softmax(mu, v, y, flows, eps)
    rows_o.sum().backward()
    rows, g_mu, g_v, dth, probs, names = _kernel(g, mu.detach(), v.detach(), y, eps, want_probs=True)
    errs = {'rows': rel_err(rows, rows_o.detach()), 'g_mu': rel_err(g_mu, mu.grad), 'g_v': rel_err(g_v, v.grad)}
    for c, lv in enumerate(leaves):
        ref = torch.stack([lv[k].grad for k in names])
        errs['dtheta[%d]' % c] = rel_err(dth[c], ref)
    with torch.no_grad():
        errs['probs'] = rel_err(probs, O.probs_mc_softmax(mu, v, flows, eps))
    record_residuals('mc_kernel:' + name, errs)
    assert all(e < TOL for e in errs.values()), errs


def _build(g):
    from tgp.pytorch_b200.dsp import config as cg
    cg.set_maximum_precission()
    cg.device = DEV
    from tgp.pytorch_b200.dsp import flows as F
    from tgp.pytorch_b200.dsp.models import instance_kernel, sparse_MF_SP
    from tgp.pytorch_b200.dsp.models.flow import instance_flow
    from tgp.pytorch_b200.dsp.likelihoods import MulticlassCategorical
    meta = g.meta
    Cn, M = meta['C'], meta['M']
    X = g.t('X')
    D = X.shape[1]
    cg.quad_points = meta['S']
    K = instance_kernel('scale_rbf', ard_num_dim=D, num_multioutput=Cn, kernel_is_shared=False,
                        init_params={'length_scale': 2.0, 'kernel_scale': 2.0, 'noisy_variance': 1e-6})
    parts = meta['flow_builder'].split(':')
    kw = {'add_f0': True} if parts[0].startswith('Step') else {}
    flows = [instance_flow(getattr(F, parts[0])(*[int(v) for v in parts[1:]], **kw)) for _ in range(Cn)]
    ip = {'variational_distribution': {'variance_scale': 1e-5, 'mean_scale': 0.0}}
    model = sparse_MF_SP(['zero', K], X, X[:M].clone(), meta['N'], MulticlassCategorical(Cn), Cn, True, False, False, False, False,
                         flows, 'single', 0.0, False, ip)
    own = dict(model.named_parameters())
    assert list(own) == list(meta['param_names']), 'parameter names / order differ from the reference'
    with torch.no_grad():
        for n in meta['param_names']:
            own[n].copy_(torch.tensor(np.asarray(g.z['param:' + n]), dtype=torch.float64).reshape(own[n].shape))
    return model.to(DEV)


@pytest.mark.parametrize('compute', ['f64', 'i8crt'])
@pytest.mark.parametrize('name', multiclass_names())
def test_model_elbo_named_gradients_and_test_nll(name, compute):
    from tgp.pytorch_b200.dsp import config as cg
    g = MulticlassGolden(name)
    model = _build(g)
    old = cg.compute
    cg.compute = compute
    try:
        X, Y = g.t('X').to(DEV), torch.tensor(np.asarray(g.z['Y'])).to(DEV)
        model.likelihood.mc_noise = g.t('eps').to(DEV)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ELBO, ELL, KLD = model.ELBO(X, Y)
            (-ELBO).backward()
        errs = {'ELBO': rel_err(ELBO.detach().cpu(), g.t('ELBO')), 'ELL': rel_err(ELL.detach().cpu(), g.t('ELL')),
                'KLD': rel_err(KLD.detach().cpu(), g.t('KLD'))}
        for n, prm in model.named_parameters():
            errs[n] = rel_err(-prm.grad.detach().cpu().reshape(-1), g.t('grad:' + n).reshape(-1))
        model.set_is_training(False)
        model.likelihood.mc_noise = g.t('eps_te').to(DEV)
        Xte, Yte = g.t('Xte').to(DEV), torch.tensor(np.asarray(g.z['Yte'])).to(DEV)
        lp, mom = model.test_log_likelihood(Xte, Yte, return_moments=True, Y_std=torch.ones(1, device=DEV), S_MC_NNet=None)
        errs['test_probs'] = rel_err(mom[0].detach().cpu().double(), g.t('test_probs'))
        record_residuals('mc_model[%s]:%s' % (compute, name), errs)
        bad = {k: e for k, e in errs.items() if not e < TOL}
        assert not bad, bad
        assert abs(float(lp) - float(g.t('test_logp'))) < 1e-5 * abs(float(g.t('test_logp')))       # scored in float32
    finally:
        cg.compute = old
        model.likelihood.mc_noise = None


def test_default_noise_is_drawn_like_rsample_and_is_seed_reproducible():
    """Without explicit noise the likelihood draws (S, C, MB) standard normals from torch's CUDA generator — what
    td.Normal(mean, std).rsample([S]) consumes on the same device — so seeded runs repeat and equal the explicit-noise call."""
    g = MulticlassGolden(multiclass_names()[0])
    model = _build(g)
    X, Y = g.t('X').to(DEV), torch.tensor(np.asarray(g.z['Y'])).to(DEV)
    torch.manual_seed(123)
    a = model.ELBO(X, Y)[0].item()
    torch.manual_seed(123)
    eps = torch.empty(g.meta['S'], g.meta['C'], X.shape[0], dtype=torch.float64, device=DEV).normal_()
    torch.manual_seed(123)
    mean = torch.zeros(g.meta['C'], X.shape[0], dtype=torch.float64, device=DEV)
    F0 = torch.distributions.Normal(mean, torch.ones_like(mean)).rsample(torch.Size([g.meta['S']]))
    assert torch.equal(F0, eps)
    model.likelihood.mc_noise = eps
    b = model.ELBO(X, Y)[0].item()
    assert a == b
