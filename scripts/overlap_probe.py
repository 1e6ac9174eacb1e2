"""prepare + qf_forward as one timed region, K_xz generation on the side stream (TGP_OPT_OVERLAP_KGEN) on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tgp.pytorch_b200 import _lib
mode = sys.argv[1] if len(sys.argv) > 1 else 'i8crt'
wl = bench.WORKLOADS['cfg4']
dev = torch.device('cuda', 0)
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(wl, 200000, gen)
p = bench.param_state(wl, X, gen)
eng, t = bench.build_engine(p, wl, dev, mode)
eng.set_params(t['Z'], t['raw_ls'], t['raw_os'], t['m'], t['L_raw'], t['log_var_noise'], t['theta'])
xb = X[:65536].to(dev)
lib = _lib.load()
for ov in (1, 0, 1, 0):
    lib.tgp_set_option(_lib.OPT_OVERLAP_KGEN, ov)
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.prepare(0.0); eng.qf_forward(xb)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(mode, 'overlap', ov, 'prepare+qf_forward median %.3f ms' % sorted(ts)[len(ts) // 2])

# the same without a host synchronisation between iterations (the host runs ahead of the device, as in a training loop)
for ov in (1, 0, 1, 0):
    lib.tgp_set_option(_lib.OPT_OVERLAP_KGEN, ov)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(20):
        eng.prepare(0.0); eng.qf_forward(xb)
    e1.record(); torch.cuda.synchronize()
    print(mode, 'overlap', ov, 'async loop: %.3f ms per (prepare + qf_forward)' % (e0.elapsed_time(e1) / 20))
