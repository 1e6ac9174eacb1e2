"""Do the tensor-core GEMMs and the ALU-bound conversion / reconstruction kernels of mode 'i8crt' overlap when they are given the
chance?  Two independent engines (own workspaces) run a forward (A) and a backward (B) of 32768 rows each, first one after the
other on one stream, then concurrently on two streams.  If the concurrent time is well below the serial sum, pipelining row chunks
over two streams inside qf_forward / qf_backward would pay.
  python scripts/concurrency_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests.gpu_util import engine_inputs, make_engine

dev = 'cuda:0'
WL = bench.WORKLOADS['cfg4']
gen = torch.Generator().manual_seed(bench.SEED)
X, Y = bench.synth(WL, 200000, gen)
p = bench.param_state(WL, X, gen)
R = 32768
engs = []
for k in range(2):
    eng, theta, _, _ = make_engine(p, WL['likelihood'], 100, dev, compute='i8crt')
    ei = engine_inputs(p, dev)
    eng.set_params(ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta)
    xb, yb = X[k * R:(k + 1) * R].to(dev), Y[k * R:(k + 1) * R].view(-1).to(dev)
    eng.prepare(0.0)
    mu, v = eng.qf_forward(xb)
    rb = eng.new_reduce_buffer()
    rows, g_mu, g_v, _ = eng.ell_forward(mu, v, yb, None, 1.0, rb)
    engs.append((eng, xb, g_mu, g_v, rb))
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def fwd(k):
    engs[k][0].qf_forward(engs[k][1])


def bwd(k):
    e, xb, g_mu, g_v, rb = engs[k]
    e.qf_backward(xb, g_mu, g_v, rb)


def timed(fn, n=8):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def concurrent(a, b):
    def run():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            a()
        with torch.cuda.stream(s2):
            b()
        cur.wait_stream(s1); cur.wait_stream(s2)
    return run


for la, a, lb, b in (('forward(A)', lambda: fwd(0), 'forward(B)', lambda: fwd(1)),
                     ('backward(A)', lambda: bwd(0), 'backward(B)', lambda: bwd(1)),
                     ('forward(A)', lambda: fwd(0), 'backward(B)', lambda: bwd(1))):
    ta, tb = timed(a), timed(b)
    tc = timed(concurrent(a, b))
    print('%-12s %.3f ms   %-12s %.3f ms   serial sum %.3f ms   concurrent on two streams %.3f ms  (%.0f %% of the sum)'
          % (la, ta, lb, tb, ta + tb, tc, 100 * tc / (ta + tb)))
