"""End-to-end: the reference's training step (loss = -ELBO; zero_grad; backward; Adam.step — trainer_base.py:329-342)
replayed through tgp.pytorch_b200.dsp on the GPU must follow the loss trajectory the unmodified reference produced
(SURVEY.md §4 test plan item 6).  25 Adam steps amplify any gradient mismatch."""
import numpy as np
import pytest
import torch

from tests.golden_util import Golden, trajectory_names, rel_err
from tests.model_util import build_from_golden

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('name', trajectory_names())
def test_adam_trajectory_matches_reference(name):
    g = Golden(name)
    model = build_from_golden(g, DEV)
    X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=g.meta['lr'])
    losses = []
    for _ in range(g.meta['steps']):
        ELBO, _, _ = model.ELBO(X, Y)
        loss = -ELBO
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.item()))
    ref = np.asarray(g.z['losses'])
    err = np.abs(np.array(losses) - ref) / np.abs(ref)
    assert err.max() < 1e-9, (err.max(), int(err.argmax()))
    for n, prm in model.named_parameters():
        assert rel_err(prm.detach().cpu().reshape(-1), g.t('final:' + n).reshape(-1)) < 1e-8, n


@pytest.mark.parametrize('name', trajectory_names())
def test_cuda_graph_step_follows_the_same_trajectory(name):
    """The captured step (tgp.pytorch_b200.train.GraphedElboStep) replayed 25 times reproduces the reference's Adam
    trajectory; warm-up steps run on a throw-away copy of the state."""
    import copy
    from tgp.pytorch_b200.train import GraphedElboStep
    g = Golden(name)
    model = build_from_golden(g, DEV)
    X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=g.meta['lr'], capturable=True)
    state0 = copy.deepcopy(model.state_dict())
    step = GraphedElboStep(model, opt, X, Y)           # warm-up + capture advance parameters and Adam moments ...
    model.load_state_dict(state0)                      # ... so restore both before the measured trajectory
    for group in opt.param_groups:
        for p in group['params']:
            st = opt.state[p]
            st['step'].zero_(); st['exp_avg'].zero_(); st['exp_avg_sq'].zero_()  # noqa: E702
    losses = []
    for _ in range(g.meta['steps']):
        losses.append(step(X, Y).clone())
    step.check()
    got = torch.stack(losses).cpu().numpy()
    ref = np.asarray(g.z['losses'])
    err = np.abs(got - ref) / np.abs(ref)
    assert err.max() < 1e-9, (err.max(), int(err.argmax()))
