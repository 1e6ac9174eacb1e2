"""Generate tests/golden/traj_*.npz: a short Adam training trajectory of the UNMODIFIED reference (O1), i.e. what
`Trainer_base.train` does per step (trainer_base.py:329-342): loss = -ELBO; zero_grad; backward; Adam.step().
The GPU test replays the same steps through tgp.pytorch_b200.dsp + torch.optim.Adam and must reproduce the losses."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import build, load_uci, randomise, flow_to_spec, OUT  # noqa: E402  (imports the reference)
from dsp.flows import SAL, StepTanhL  # noqa: E402
from dsp.models.flow import CompositeFlow  # noqa: E402


def run(name, kind, flow, steps=25, lr=0.01):
    Xb, Yb, Xbt, Ybt, ys = load_uci('boston')
    model = build(kind, Xb, 100, float(Xb.shape[0]), flow, seed=31)
    randomise(model, 77)
    store = {'X': Xb.numpy(), 'Y': Yb.numpy(), 'Xte': Xbt.numpy(), 'Yte': Ybt.numpy()}
    names = []
    for n, prm in model.named_parameters():
        store['param:' + n] = prm.detach().numpy().copy()
        names.append(n)
    with torch.no_grad():
        fl = model.G_matrix[0]
        comp = fl if isinstance(fl, CompositeFlow) else CompositeFlow([fl])
        spec = flow_to_spec(comp, store, Xb, 'fl')
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    losses = []
    for _ in range(steps):
        E, ELL, KLD = model.ELBO(Xb, Yb)
        loss = -E
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    store['losses'] = np.array(losses)
    for n, prm in model.named_parameters():
        store['final:' + n] = prm.detach().numpy().copy()
    store['meta'] = np.array(json.dumps({'name': name, 'likelihood': 'gauss_linear' if kind == 'SVGP' else 'gauss_nonlinear',
                                          'N': float(Xb.shape[0]), 'M': 100, 'y_std': ys, 'n_quad': 100, 'flow_train': spec,
                                          'flow_test': spec, 'param_names': names, 'id_flow': False, 'steps': steps, 'lr': lr}))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **store)
    print(name, losses[0], losses[-1])


if __name__ == '__main__':
    run('traj_boston_tgp_steptanh13', 'TGP', StepTanhL(1, 3, add_f0=True))
    run('traj_boston_svgp', 'SVGP', None)
