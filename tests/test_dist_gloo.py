"""N > 1 host logic on CPU: world-size-2 gloo processes shard the rows of one global minibatch, evaluate their slice
with the GLOBAL scale (the oracle stands in for the kernels — there is no GPU here), all-reduce ONE packed buffer and
must reproduce the single-process ELBO and gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import tgp_oracle as O
    from tests.golden_util import Golden
    from tgp.pytorch_b200 import dist as D
    g = Golden('synth_reg_d8_m64_p1')
    p = g.oracle_params('train')
    X, Y = g.t('X'), g.t('Y').view(-1)
    sl = D.local_slice(X.shape[0], rank, world)
    scale = D.global_scale(g.meta['N'], X.shape[0])
    leaves = O.leaf_params(p)
    for t in leaves.values():
        t.requires_grad_(True)
    mu, v = O.qf_marginals(X[sl], p)
    ell = scale * O.ell_rows(mu, v, Y[sl], p, g.meta['likelihood'], g.meta['n_quad']).sum()
    grads = torch.autograd.grad(ell, list(leaves.values()), allow_unused=True)
    grads = [torch.zeros_like(t) if gr is None else gr for t, gr in zip(leaves.values(), grads)]
    like = [ell.detach().reshape(1)] + grads
    buf = torch.cat([t.reshape(-1) for t in like])
    dist.all_reduce(buf)                                    # the one collective of the step
    parts, o = [], 0
    for t in like:
        parts.append(buf[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    if rank == 0:
        ret['ell'] = parts[0].clone()
        ret['grads'] = {k: t.clone() for k, t in zip(leaves.keys(), parts[1:])}
    dist.destroy_process_group()


def test_two_rank_row_sharding_matches_single_process():
    from oracle import tgp_oracle as O
    from tests.golden_util import Golden, rel_err
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29533, ret), nprocs=world, join=True)
    g = Golden('synth_reg_d8_m64_p1')
    p = g.oracle_params('train')
    E, ELL, KLD, rows, grads = O.elbo_and_grads(g.t('X'), g.t('Y').view(-1), p, g.meta['N'], g.meta['likelihood'],
                                                g.meta['n_quad'])
    assert rel_err(ret['ell'], ELL) < 1e-13
    # ELBO gradients = ELL gradients (all-reduced) + KL gradients (replicated, rank-agnostic)
    leaves = O.leaf_params(p)
    for t in leaves.values():
        t.grad = None
    kl_grads = torch.autograd.grad(O.kl_whitened(p), [leaves['m'], leaves['L_raw']])
    total = dict(ret['grads'])
    total['m'] = total['m'] - kl_grads[0]
    total['L_raw'] = total['L_raw'] - kl_grads[1]
    for k, ref in grads.items():
        assert rel_err(total[k], ref) < 1e-11, k


def test_shard_bounds_cover_ragged_batches():
    from tgp.pytorch_b200 import dist as D
    for n in (0, 1, 7, 455, 65536):
        for w in (1, 2, 3, 8):
            b = D.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(w))
            assert max(b[i + 1] - b[i] for i in range(w)) - min(b[i + 1] - b[i] for i in range(w)) <= 1


def _mlp_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tgp.pytorch_b200 import dist as D
    from tgp.pytorch_b200 import functional as Fn
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.Tanh(), torch.nn.Linear(6, 1)).double()
    X = torch.randn(37, 4, dtype=torch.float64)
    w = torch.randn(37, dtype=torch.float64)
    sl = D.local_slice(37, rank, world)
    out = Fn.synced_module_call(net, X[sl]).squeeze(-1)          # product path: rank-local rows, gradients summed over ranks
    (out * w[sl]).sum().backward()
    with Fn.local_only():                                        # yardstick: the whole batch on this rank, no collective
        ref = [p.grad.clone() for p in net.parameters()]
        for p in net.parameters():
            p.grad = None
        (Fn.synced_module_call(net, X).squeeze(-1) * w).sum().backward()
        full = [p.grad.clone() for p in net.parameters()]
    ret[rank] = max(float((a - b).abs().max()) for a, b in zip(ref, full))
    dist.destroy_process_group()


def test_flow_mlp_gradients_are_summed_over_ranks():
    """ID_TGP under row sharding: the flow MLPs see rank-local rows; `functional.synced_module_call` must
    deliver the global-batch gradient to every rank."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_mlp_worker, args=(world, 29547, ret), nprocs=world, join=True)
    assert all(ret[r] < 1e-13 for r in range(world)), dict(ret)


def _std_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tgp.pytorch_b200 import dist as D
    from tgp.pytorch_b200 import functional as Fn
    g = torch.Generator().manual_seed(11)
    v = torch.rand(1, 1001, generator=g, dtype=torch.float64) * 3.0 + 1e3          # (Dy, MB) variances with a large common offset
    sl = D.local_slice(1001, rank, world)
    got = Fn.global_std(v[:, sl])
    with Fn.local_only():
        alone = Fn.global_std(v)
    ret[rank] = (float(abs(got - v.std()) / v.std()), float(abs(alone - v.std())))
    dist.destroy_process_group()


def test_batch_wide_std_of_the_bernoulli_moments_spans_all_ranks():
    """SURVEY.md §8e exception: the reference's Bernoulli `marginal_moments` uses `gauss_cov.std()` of the whole batch
    (Bernoulli.py:120,141); row-sharded evaluation must see the same number on every rank."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_std_worker, args=(world, 29549, ret), nprocs=world, join=True)
    assert all(ret[r][0] < 1e-13 and ret[r][1] == 0.0 for r in range(world)), dict(ret)
