"""What bounds a step when a rank holds few rows (strong scaling: 65536 / 8 = 8192 rows per GPU)?  One GPU, cfg4 parameters,
ROWS rows per step: device time per step (CUDA events), host time to ENQUEUE a step (wall clock without synchronising), the same
with the factorisation kept on the caller's stream, and the stage times.
  python scripts/small_step.py [rows=8192] [compute=i8crt]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tgp.pytorch_b200 import _lib, functional as Fn

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
mode = sys.argv[2] if len(sys.argv) > 2 else 'i8crt'
wl = bench.WORKLOADS['cfg4']
dev = torch.device('cuda', 0)
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(wl, 200000, gen)
p = bench.param_state(wl, X, gen)
eng, t = bench.build_engine(p, wl, dev, mode)
leaves = [t['Z'], t['raw_ls'], t['raw_os'], t['m'], t['L_raw'], t['log_var_noise'], t['theta']]
for x in leaves:
    if x is not None:
        x.requires_grad_(True)
xb, yb = X[:rows].to(dev), Y[:rows].view(-1).to(dev)
lib = _lib.load()


def step(check):
    for x in leaves:
        if x is not None:
            x.grad = None
    ELL, KLD, _, _, _ = Fn.elbo_terms(eng, xb, yb, wl['N'] / 65536, *leaves, None, check_status=check, sync_ell=False)
    (-(ELL - KLD)).backward()


def measure(label, check, n=30):
    for _ in range(5):
        step(check)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        step(check)
    e1.record(); t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    print('%-46s device %.3f ms/step   host enqueue %.3f ms/step' % (label, e0.elapsed_time(e1) / n, 1e3 * t_enq / n))


measure('%s rows=%d, status check on' % (mode, rows), True)
measure('%s rows=%d, status check off' % (mode, rows), False)
lib.tgp_set_option(_lib.OPT_OVERLAP_KGEN, 0)
measure('... factorisation on the caller stream', False)
lib.tgp_set_option(_lib.OPT_OVERLAP_KGEN, 1)

# the same step as ONE CUDA graph (no host launch cost)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step(False)
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
for x in leaves:
    if x is not None:
        x.grad = None
with torch.cuda.graph(g):
    step(False)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    g.replay()
e1.record(); torch.cuda.synchronize()
print('%-46s device %.3f ms/step' % ('... replayed as one CUDA graph', e0.elapsed_time(e1) / 30))
