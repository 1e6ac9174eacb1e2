"""Step latency at the reference's own problem sizes (boston 455x13, power 2048x4; M = 100): eager step, CUDA-graph step with
torch's capturable Adam, CUDA-graph step with FusedAdam (one launch for all parameter tensors)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.golden_util import Golden
from tests.model_util import build_from_golden
from tgp.pytorch_b200.train import GraphedElboStep, FusedAdam
warnings.simplefilter('ignore')
dev = 'cuda:0'
for name in ('boston_svgp_p1', 'boston_tgp_steptanh13_p1', 'boston_tgp_steptanh102_p1', 'power_tgp_sal2_p1', 'boston_idtgp_nodrop_p1'):
    g = Golden(name)
    model = build_from_golden(g, dev)
    X, Y = g.t('X').to(dev), g.t('Y').to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    def eager():
        ELBO, _, _ = model.ELBO(X, Y); loss = -ELBO
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
        return loss
    for _ in range(5): eager()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): l = eager()
    l.item(); t_eager = (time.perf_counter() - t0) / 50
    # a fresh model: autograd ties each parameter's gradient accumulator to the stream of its first backward, and a
    # capture must not depend on the legacy default stream the eager steps above ran on
    model = build_from_golden(g, dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    step = GraphedElboStep(model, opt, X, Y)
    for _ in range(5): step(X, Y)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): l = step(X, Y)
    l.item(); t_graph = (time.perf_counter() - t0) / 200
    step.check()
    model = build_from_golden(g, dev)
    step = GraphedElboStep(model, FusedAdam(model.parameters(), lr=1e-3), X, Y)
    for _ in range(5): step(X, Y)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): l = step(X, Y)
    l.item(); t_fused = (time.perf_counter() - t0) / 200
    step.check()
    print('%-28s rows %5d  eager %.3f ms/step  graph %.3f ms/step  graph + FusedAdam %.3f ms/step  (%.0f rows/s)'
          % (name, X.shape[0], t_eager * 1e3, t_graph * 1e3, t_fused * 1e3, X.shape[0] / t_fused))
