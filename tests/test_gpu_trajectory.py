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


@pytest.mark.parametrize('name', trajectory_names())
@pytest.mark.parametrize('graphed', [False, True])
def test_fused_adam_follows_the_reference_trajectory(name, graphed):
    """tgp.pytorch_b200.train.FusedAdam (one launch for all parameter tensors, step count on the device) in the reference's
    training step, eager and captured as one CUDA graph: the reference's 25-step loss trajectory and final parameters."""
    import copy
    from tgp.pytorch_b200.train import FusedAdam, GraphedElboStep
    g = Golden(name)
    model = build_from_golden(g, DEV)
    X, Y = g.t('X').to(DEV), g.t('Y').to(DEV)
    opt = FusedAdam(model.parameters(), lr=g.meta['lr'])
    losses = []
    if graphed:
        state0 = copy.deepcopy(model.state_dict())
        step = GraphedElboStep(model, opt, X, Y)
        model.load_state_dict(state0)
        for st in opt.state.values():
            st['exp_avg'].zero_(); st['exp_avg_sq'].zero_()  # noqa: E702
        opt.step_dev.zero_()
        for _ in range(g.meta['steps']):
            losses.append(float(step(X, Y).item()))
        step.check()
    else:
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


def test_fused_adam_matches_torch_adam_with_weight_decay_groups():
    from tgp.pytorch_b200.train import FusedAdam
    gen = torch.Generator().manual_seed(0)
    shapes = [(5000,), (3, 7), (1,), ()]
    a = [torch.randn(s, generator=gen, dtype=torch.float64).to(DEV).requires_grad_(True) for s in shapes]
    b = [t.detach().clone().requires_grad_(True) for t in a]
    groups = lambda ps: [{'params': ps[:2], 'lr': 0.01}, {'params': ps[2:], 'lr': 0.003, 'weight_decay': 1e-5}]  # noqa: E731
    o1, o2 = torch.optim.Adam(groups(a)), FusedAdam(groups(b))
    for it in range(7):
        for ps, o in ((a, o1), (b, o2)):
            o.zero_grad()
            loss = sum(((p - 0.3 * (i + 1)) ** 2).sum() * (it + 1) for i, p in enumerate(ps))
            loss.backward()
            o.step()
    for x, y in zip(a, b):
        assert rel_err(y.detach().cpu(), x.detach().cpu()) < 1e-13
