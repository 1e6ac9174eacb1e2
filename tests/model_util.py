"""Builds `tgp.pytorch_b200.dsp` models the way the reference's main.py does and loads golden parameters by NAME
(the module tree reproduces the reference's parameter names, which main.py:281-286 greps)."""
import numpy as np
import torch

from tests.golden_util import Golden


def build_from_golden(g, device):
    from tgp.pytorch_b200.dsp import config as cg
    cg.set_maximum_precission()
    cg.device = device
    from tgp.pytorch_b200.dsp.models import instance_kernel, sparse_MF_SP, sparse_MF_GP
    from tgp.pytorch_b200.dsp.models.flow import instance_flow
    from tgp.pytorch_b200.dsp.likelihoods import GaussianNonLinearMean, GaussianLinearMean, Bernoulli
    from tgp.pytorch_b200.dsp.flows import SAL, StepTanhL
    meta = g.meta
    X = g.t('X')
    M, D = meta['M'], X.shape[1]
    K = instance_kernel('scale_rbf', ard_num_dim=D, num_multioutput=1, kernel_is_shared=False,
                        init_params={'length_scale': 2.0, 'kernel_scale': 2.0, 'noisy_variance': 1e-6})
    Z = torch.tensor(np.asarray(g.z['param:Z']), dtype=torch.float64)[0].clone()      # placeholder of the right shape
    ip = {'variational_distribution': {'variance_scale': 1e-5, 'mean_scale': 0.0}}
    lik_kind = meta['likelihood']
    if lik_kind == 'gauss_linear':
        lik = GaussianLinearMean(out_dim=1, noise_init=0.05, noise_is_shared=False)
        model = sparse_MF_GP(['zero', K], X, Z, meta['N'], lik, 1, True, False, False, False, False, 0.0, ip)
    else:
        # rebuild the flow architecture from the fixture's layer list
        spec = meta['flow_train']
        names = meta['param_names']
        if meta.get('flow_builder'):
            from tgp.pytorch_b200.dsp.flows import ArcSL, BoxCoxL, build_chain
            parts = meta['flow_builder'].split(':')
            constraint = (lambda lam: 2.0 * torch.sigmoid(lam) + 0.01) if meta.get('boxcox_constraint') else None
            if parts[0].startswith('Step'):
                from tgp.pytorch_b200.dsp import flows as F
                kw = {'add_f0': True} if parts[0] in ('StepSAL', 'StepInverseBoxCoxL') else {}
                flow = getattr(F, parts[0])(*[int(v) for v in parts[1:]], **kw)
            elif parts[0] == 'ArcSL':
                flow = ArcSL(int(parts[1]))
            elif parts[0] == 'BoxCoxL':
                flow = BoxCoxL(int(parts[1]), init_random=True)
            else:
                flow = build_chain(parts[1], int(parts[2]), constraint=constraint)
        elif spec[0][0] == 'tanh_step':
            flow = StepTanhL(len(spec) // 2, len(spec[0][1]), add_f0=True)
        elif meta['id_flow']:
            nb = len(spec) // 2
            first = [n for n in names if 'NNets_a.0.forward_lin.0.weight' in n][0]
            H = g.z['param:' + first].shape[0]
            n_hidden = len([n for n in names if 'flow_arr.0.NNets_a' in n and n.endswith('weight')]) - 1
            act = 'tanh' if D == 13 else 'relu'
            flow = instance_flow(SAL(nb, input_dependent=True, input_dim=D, inference='MC_dropout', hidden_activation=act,
                                     num_hidden_layers=n_hidden, dropout=0.5 if D == 13 else 0.25, batch_norm=0,
                                     hidden_dim=H))
            flow.turn_off_initializer_parameters()
        else:
            flow = SAL(len(spec) // 2)
        lik = Bernoulli() if lik_kind == 'bernoulli' else \
            GaussianNonLinearMean(out_dim=1, noise_init=0.05, noise_is_shared=False, quadrature_points=meta['n_quad'])
        model = sparse_MF_SP(['zero', K], X, Z, meta['N'], lik, 1, True, False, False, False, False, [flow], 'single',
                             0.0, False, ip)
    own = dict(model.named_parameters())
    missing = [n for n in meta['param_names'] if n not in own]
    extra = [n for n in own if n not in meta['param_names']]
    assert not missing and not extra, ('parameter names differ from the reference', missing, extra)
    with torch.no_grad():
        for n in meta['param_names']:
            own[n].copy_(torch.tensor(np.asarray(g.z['param:' + n]), dtype=torch.float64).reshape(own[n].shape))
    return model.to(device)


def set_dropout_mode(model, g):
    """Fixtures recorded with dropout OFF: put the dropout layers in eval mode.  Fixtures recorded in TRAINING mode
    (meta['dropout']): keep them active and hand every input-dependent layer the reference's recorded keep-masks."""
    if not g.meta['id_flow']:
        return
    if g.meta.get('dropout'):
        for li, layer in enumerate(model.G_matrix[0].flow_arr):
            key = 'dropmask:%d' % li
            if key in g.z.files:
                layer.dropout_masks = torch.tensor(np.asarray(g.z[key]), dtype=torch.uint8)
        return
    for m in model.modules():
        if 'Dropout' in type(m).__name__:
            m.eval()
