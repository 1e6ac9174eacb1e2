"""Bridges oracle-style parameter dicts (tests/golden_util.Golden.oracle_params) to the CUDA engine."""
import torch

from tgp.pytorch_b200.engine import Engine, FlowLayout


def flow_layout_and_params(layers, device):
    """oracle layer list -> (FlowLayout, theta tensor, rowparams tensor or None, leaf-name list for theta)."""
    desc, theta, rowcols, names = [], [], [], []
    flat = []                                       # (layer, leaf-name prefix, switch_off pair or None)
    for i, lay in enumerate(layers):
        if lay[0] == 'step_group':
            flat.append((lay, 'flow%d' % i, None))
            flat += [(m, 'flow%d.m%d' % (i, j), sw) for j, (m, sw) in enumerate(lay[1])]
        else:
            flat.append((lay, 'flow%d' % i, None))
    for lay, pre, sw in flat:
        n_desc = len(desc)
        i = pre[4:]                                 # names below are 'flow%s...' % i
        if lay[0] == 'identity':
            continue
        if lay[0] == 'step_group':
            desc.append(dict(kind='step_group', n_steps=len(lay[1]), add_f0=lay[2]))
        elif lay[0] == 'affine':
            desc.append(dict(kind='affine', restrict=lay[3]))
            theta += [lay[1].reshape(()), lay[2].reshape(())]
            names += ['flow%s.a' % i, 'flow%s.b' % i]
        elif lay[0] == 'tanh_step':
            desc.append(dict(kind='tanh_step', n_steps=len(lay[1]), add_f0=lay[2]))
            for j, st in enumerate(lay[1]):
                theta += [t.reshape(()) for t in st]
                names += ['flow%s.%d.%s' % (i, j, c) for c in 'abcd']
        elif lay[0] == 'arcsinh':
            desc.append(dict(kind='arcsinh', restrict=lay[5], add_f0=lay[6]))
            theta += [t.reshape(()) for t in lay[1:5]]
            names += ['flow%s.%s' % (i, c) for c in 'abcd']
        elif lay[0] in ('boxcox', 'invboxcox'):
            desc.append(dict(kind=lay[0], add_f0=lay[2]))
            theta += [lay[1].reshape(())]
            names += ['flow%s.lam' % i]
        elif lay[0] == 'sal':
            per_row = lay[1].dim() > 0
            desc.append(dict(kind='sal', restrict=lay[3], add_f0=lay[4], per_row=per_row))
            if per_row:
                rowcols += [lay[1], lay[2]]
            else:
                theta += [lay[1].reshape(()), lay[2].reshape(())]
                names += ['flow%s.a' % i, 'flow%s.b' % i]
        if sw is not None:
            assert len(desc) == n_desc + 1
            desc[-1]['switch'] = True
            theta += [sw[0].reshape(()), sw[1].reshape(())]
            names += ['flow%s.sw_a' % i, 'flow%s.sw_b' % i]
    fl = FlowLayout(desc)
    th = torch.stack(theta).to(device) if theta else torch.zeros(0, dtype=torch.float64, device=device)
    rp = torch.stack(rowcols, dim=1).contiguous().to(device) if rowcols else None
    return fl, th.double().contiguous(), rp, names


def engine_inputs(p, device):
    dev = torch.device(device)
    f = lambda t: t.detach().to(dev).double().contiguous()  # noqa: E731
    return dict(Z=f(p['Z']), raw_ls=f(p['raw_lengthscale'].reshape(-1)), raw_os=f(p['raw_outputscale'].reshape(1)),
                m=f(p['m']), L_raw=f(p['L_raw']), log_var_noise=f(p['log_var_noise'].reshape(1)))


def make_engine(p, likelihood, n_quad, device, compute='f64'):
    fl, theta, rowp, names = flow_layout_and_params(p['flow'], device)
    if likelihood == 'gauss_linear':
        fl = FlowLayout([])
        theta = torch.zeros(0, dtype=torch.float64, device=device)
    eng = Engine(p['Z'].shape[0], p['Z'].shape[1], likelihood, n_quad, fl, device, compute=compute)
    return eng, theta, rowp, names
