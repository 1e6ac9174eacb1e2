"""Runs `n` complete cfg4 steps (prepare, forward, epilogue, backward, chain) through the engine — the target of the ncu
launch-list captures (`ncu --metrics gpu__time_duration.sum --clock-control none python scripts/one_step.py f64 2`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests.gpu_util import engine_inputs, make_engine
dev = 'cuda:0'
compute = sys.argv[1] if len(sys.argv) > 1 else 'f64'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(200000, bench.D, gen)
p = bench.param_state(X, gen)
xb, yb = X[:65536].to(dev), Y[:65536].view(-1).to(dev)
eng, theta, _, _ = make_engine(p, 'gauss_nonlinear', 100, dev, compute=compute)
ei = engine_inputs(p, dev)
eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
for _ in range(n):
    eng.prepare(0.0)
    mu, v = eng.qf_forward(xb)
    rb = eng.new_reduce_buffer()
    rows, g_mu, g_v, _ = eng.ell_forward(mu, v, yb, None, 5e6 / 65536, rb)
    eng.qf_backward(xb, g_mu, g_v, rb)
    eng.chain_backward(rb, 1.0, -1.0)
torch.cuda.synchronize()
print('ok')
