"""Device time of tgp_prepare alone (median of 20) — the per-step factorisation latency."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests.gpu_util import engine_inputs, make_engine
dev = 'cuda:0'
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(200000, bench.D, gen)
p = bench.param_state(X, gen)
eng, theta, _, _ = make_engine(p, 'gauss_nonlinear', 100, dev, compute='f64')
ei = engine_inputs(p, dev)
eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
ts = []
for i in range(25):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); eng.prepare(0.0); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print('prepare median %.3f ms  min %.3f ms' % (sorted(ts[5:])[10], min(ts[5:])))
