"""The input-dependent flow MLPs of ID_TGP on the device (tgp_flow_mlp_forward / _backward, csrc/flow_mlp.cuh): against the
torch modules, against the oracle with explicit dropout masks, the Philox stream's statistics and its exported mask."""
import pytest
import torch
import torch.nn as nn

from oracle import tgp_oracle as O
from tests.golden_util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _nets(n_nets, n_in, H, L, act, p, seed):
    from tgp.pytorch_b200.dsp.nn_layers import apply_linear
    torch.manual_seed(seed)
    nets = []
    for _ in range(n_nets):
        layers, d = [], n_in
        for _l in range(L):
            layers.append(apply_linear(d, H, act, shape=None, std=0.0, drop=p, bn=0))
            d = H
        layers.append(apply_linear(H, 1, 'linear', shape=None, std=0.0, drop=0.0, bn=0))
        nets.append(nn.Sequential(*layers).double().to(DEV))
    return nets


def _weights_cpu(net):
    return [(b.forward_lin[0].weight.detach().cpu(), b.forward_lin[0].bias.detach().cpu()) for b in net]


@pytest.mark.parametrize('n_nets,n_in,H,L,act', [(2, 13, 25, 1, 'tanh'), (2, 4, 50, 2, 'relu'), (4, 7, 64, 3, 'sigmoid'), (1, 1, 1, 1, 'relu'),
                                                 (3, 64, 33, 4, 'tanh')])
def test_kernel_matches_torch_modules_without_dropout(n_nets, n_in, H, L, act):
    from tgp.pytorch_b200 import functional as Fn
    nets = _nets(n_nets, n_in, H, L, act, 0.0, seed=n_in + H)
    spec = Fn.mlp_spec(nets)
    assert spec is not None and spec['L'] == L and spec['H'] == H
    g = torch.Generator().manual_seed(1)
    X = torch.randn(301, n_in, generator=g, dtype=torch.float64).to(DEV)
    w = torch.randn(301, n_nets, generator=g, dtype=torch.float64).to(DEV)
    out = Fn.flow_mlp(None, nets, spec, X)
    (out * w).sum().backward()
    got = [p.grad.clone() for net in nets for p in net.parameters()]
    for net in nets:
        net.zero_grad()
    ref = torch.stack([net(X).squeeze(-1) for net in nets], dim=1)
    (ref * w).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach().cpu()) < 1e-13
    for a, p in zip(got, [p for net in nets for p in net.parameters()]):
        assert rel_err(a.cpu(), p.grad.cpu()) < 1e-12


def test_explicit_masks_match_the_oracle_and_its_autograd():
    from tgp.pytorch_b200 import functional as Fn
    n_nets, n_in, H, L, p = 2, 4, 50, 2, 0.25
    nets = _nets(n_nets, n_in, H, L, 'relu', p, seed=3)
    spec = Fn.mlp_spec(nets)
    assert spec['training'] and spec['p'] == p
    g = torch.Generator().manual_seed(2)
    R = 257
    X = torch.randn(R, n_in, generator=g, dtype=torch.float64)
    w = torch.randn(R, n_nets, generator=g, dtype=torch.float64)
    masks = (torch.rand(n_nets, L, R, H, generator=g) > p).to(torch.uint8)
    out = Fn.flow_mlp(None, nets, spec, X.to(DEV), masks)
    (out * w.to(DEV)).sum().backward()
    for ni, net in enumerate(nets):
        ws = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in _weights_cpu(net)]
        ref = O.flow_mlp(X, ws, 'relu', p, [masks[ni, l] for l in range(L)])
        (ref * w[:, ni]).sum().backward()
        assert rel_err(out[:, ni].detach().cpu(), ref.detach()) < 1e-13
        for (W, b), blk in zip(ws, net):
            assert rel_err(blk.forward_lin[0].weight.grad.cpu(), W.grad) < 1e-12
            assert rel_err(blk.forward_lin[0].bias.grad.cpu(), b.grad) < 1e-12


def test_philox_dropout_statistics_export_and_replay():
    from tgp.pytorch_b200 import functional as Fn

    class Owner:
        last_dropout_masks = None
    n_nets, n_in, H, L, p = 2, 8, 48, 2, 0.25
    nets = _nets(n_nets, n_in, H, L, 'tanh', p, seed=4)
    spec = Fn.mlp_spec(nets)
    X = torch.randn(4096, n_in, dtype=torch.float64, device=DEV)
    own = Owner()
    out1 = Fn.flow_mlp(own, nets, spec, X).detach()
    m1 = own.last_dropout_masks.clone()
    out2 = Fn.flow_mlp(own, nets, spec, X).detach()
    m2 = own.last_dropout_masks.clone()
    assert m1.shape == (n_nets, L, 4096, H)
    n = m1.numel()
    keep = float(m1.float().mean())
    assert abs(keep - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5                       # keep rate 1 - p within 5 sigma
    assert 0.3 < float((m1 != m2).float().mean()) / (2 * p * (1 - p)) < 1.7         # a fresh stream position per call
    # no structure along rows / units: column and row keep rates are all near 1 - p
    assert float((m1.float().mean(dim=2) - (1 - p)).abs().max()) < 6 * (p * (1 - p) / 4096) ** 0.5
    # the exported mask replayed as an explicit one reproduces the output bit for bit
    out1r = Fn.flow_mlp(own, nets, spec, X, m1).detach()
    assert torch.equal(out1, out1r) and not torch.equal(out1, out2)
    # eval mode: dropout off, deterministic
    for net in nets:
        net.eval()
    spec_e = Fn.mlp_spec(nets)
    assert not spec_e['training']
    a = Fn.flow_mlp(own, nets, spec_e, X).detach()
    assert own.last_dropout_masks is None
    assert rel_err(a.cpu(), torch.stack([net(X).squeeze(-1) for net in nets], dim=1).detach().cpu()) < 1e-13


def test_unsupported_architectures_fall_back_to_the_modules_not_to_wrong_numbers():
    from tgp.pytorch_b200 import functional as Fn
    from tgp.pytorch_b200.dsp.nn_layers import apply_linear
    wide = [nn.Sequential(apply_linear(4, 80, 'relu'), apply_linear(80, 1, 'linear'))]
    bn = [nn.Sequential(apply_linear(4, 8, 'relu', bn=1), apply_linear(8, 1, 'linear'))]
    assert Fn.mlp_spec(wide) is None and Fn.mlp_spec(bn) is None
