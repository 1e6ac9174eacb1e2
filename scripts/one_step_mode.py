"""One warm-up + one measured ELBO forward+backward step at a bench workload in a given compute mode (for ncu launch lists).
  python scripts/one_step_mode.py i8crt [cfg4|cfg5]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tgp.pytorch_b200 import functional as Fn
mode = sys.argv[1] if len(sys.argv) > 1 else 'f64'
wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else 'cfg4']
dev = torch.device('cuda', 0)
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(wl, 200000, gen)
p = bench.param_state(wl, X, gen)
eng, t = bench.build_engine(p, wl, dev, mode)
leaves = [t['Z'], t['raw_ls'], t['raw_os'], t['m'], t['L_raw'], t['log_var_noise'], t['theta']]
for x in leaves:
    if x is not None:
        x.requires_grad_(True)
xb, yb = X[:65536].to(dev), Y[:65536].view(-1).to(dev)
for it in range(2):
    for x in leaves:
        if x is not None:
            x.grad = None
    torch.cuda.synchronize()
    if it == 1:
        torch.zeros(1, device=dev).add_(1)          # marker launch: the step of interest follows
    ELL, KLD, _, _, _ = Fn.elbo_terms(eng, xb, yb, wl['N'] / 65536, *leaves, None, check_status=True)
    (-(ELL - KLD)).backward()
    torch.cuda.synchronize()
print('loss', float((ELL - KLD).item()))
