#!/usr/bin/env python
"""bench.py — minibatch ELBO forward+backward throughput of the B200 path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one process per GPU under torchrun
  python bench.py --impl reference --gpus N ...            # the reference's algorithm on the host CPU cores (oracle port)

Workload (BASELINE.json configs[3]): TGP regression, synthetic N = 5 M rows, D = 8, M = 1024 inducing
points, StepTanhL(1,3) flow, Gaussian likelihood, 100 Gauss-Hermite points, FP64 (what the reference's main.py runs);
one step = ELBO forward + backward over one minibatch of 65536 rows PER GPU (weak scaling; the global minibatch is
65536 * N rows, N/MB scaling uses the global size, one NCCL all-reduce of the packed pre-chain gradient buffer).

One JSON line on stdout (rank 0).  Keys per the driver contract plus `roofline` and `cpu_baseline`.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DATA, D, M, BATCH = 5_000_000, 8, 1024, 65536
N_QUAD = 100
SEED = 1234
CPU_SAMPLE_ROWS = 8192           # rows per CPU step: a bounded sample of the same workload
METRIC = 'elbo_fwd_bwd_rows_per_s'


def synth(n, d, gen):
    X = torch.randn(n, d, generator=gen, dtype=torch.float64)
    w = torch.randn(d, generator=gen, dtype=torch.float64)
    y = torch.sinh(0.7 * (X @ w) / math.sqrt(d)) + 0.1 * torch.randn(n, generator=gen, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    return X, y.view(-1, 1)


def param_state(X, gen):
    """'Mid-training' state P1: nothing at a trivial value; K_zz stays well conditioned."""
    idx = torch.randperm(X.shape[0], generator=gen)[:M]
    inv_sp = lambda t: t + torch.log(-torch.expm1(-t))  # noqa: E731
    p = dict(Z=X[idx].clone(),
             raw_lengthscale=inv_sp(1.5 + 1.0 * torch.rand(D, generator=gen, dtype=torch.float64)),
             raw_outputscale=inv_sp(torch.tensor(1.5, dtype=torch.float64)),
             m=torch.randn(M, generator=gen, dtype=torch.float64),
             L_raw=0.5 * torch.eye(M, dtype=torch.float64) + 0.05 * torch.randn(M, M, generator=gen, dtype=torch.float64),
             log_var_noise=torch.log(torch.tensor(0.05, dtype=torch.float64)))
    steps = []
    for _ in range(3):
        e = torch.randn(4, generator=gen, dtype=torch.float64)
        steps.append((e[0].clone(), inv_sp(torch.abs((e[1] + 1.0) / 3.0) + 1e-3), e[2].clone(),
                      inv_sp(torch.abs((e[3] + 1.0) / 3.0) + 1e-3)))
    p['flow'] = [('tanh_step', steps, True), ('affine', torch.tensor(1.1, dtype=torch.float64),
                                               torch.tensor(-0.05, dtype=torch.float64), False)]
    return p


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML from a background thread every 100 ms;
    the same fields as the nvidia-smi query of the profiling recipe, without forking nvidia-smi under the load)."""

    def __init__(self, gpu_index):
        self.gpu, self.rows, self._stop, self._thr, self._nvml = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[self.gpu]) if vis and vis.split(',')[self.gpu].isdigit() else self.gpu
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.rows.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        if self._nvml is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable']}
        self._stop = True
        self._thr.join(timeout=1.0)
        nv = self._nvml
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        reasons = set()
        for _, _, rs in self.rows:
            for n, bit in names.items():
                if rs & bit:
                    reasons.add(n)
        sm = [r[0] for r in self.rows]
        mx = [r[1] for r in self.rows]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measure_fp64_peak(dev):
    """cuBLAS DGEMM 8192^3 via torch.matmul, best of 5 — MEASURED_PEAKS.json has no FP64 figure (method of that file)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()  # noqa: E702
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(p, X, Y, rows, steps, warmup):
    """ELBO + backward of the reference's algorithm (oracle port O2, pinned to the reference by tests/golden) on the
    host cores; each step is a `rows`-row sample of the workload."""
    from oracle import tgp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    times = []
    for s in range(warmup + steps):
        lo = (s * rows) % (X.shape[0] - rows)
        xb, yb = X[lo:lo + rows], Y[lo:lo + rows].view(-1)
        t0 = time.perf_counter()
        O.elbo_and_grads(xb, yb, p, float(N_DATA), 'gauss_nonlinear', N_QUAD)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return rows / float(np.median(times)), float(np.median(times))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    gen = torch.Generator().manual_seed(SEED)
    n_cpu = max(CPU_SAMPLE_ROWS * 8, 200_000)
    X, Y = synth(n_cpu, D, gen)
    p = param_state(X, gen)
    rows_s, sec = cpu_reference_steps(p, X, Y, CPU_SAMPLE_ROWS, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {'impl': 'reference', 'metric': METRIC, 'value': rows_s, 'unit': 'rows/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': rows_s, 'unit': 'rows/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d-row minibatch per step of the same workload (M=%d, D=%d, %d GH points, FP64), '
                                       'torch CPU with %d threads' % (CPU_SAMPLE_ROWS, M, D, N_QUAD, cores)},
            'e2e': {'value': rows_s, 'unit': 'rows/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


def workload_config(n_gpus):
    return {'workload': 'TGP regression synthetic N=5M D=8 M=1024 batch 65536/GPU, StepTanhL(1,3), Gaussian lik, '
                        '100 GH points (BASELINE configs[3])', 'N': N_DATA, 'D': D, 'M': M, 'rows_per_gpu_per_step': BATCH,
            'global_batch': BATCH * n_gpus, 'parallelism': 'rows x%d' % n_gpus,
            'l2': 'per-step working set (A|B workspace 1.07 GB) >> 126 MB L2, plus an explicit 256 MiB L2 flush between steps'}


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from tests.gpu_util import engine_inputs, make_engine
    from tgp.pytorch_b200 import _lib, functional as Fn
    lib = _lib.load()

    gen = torch.Generator().manual_seed(SEED)
    X, Y = synth(N_DATA, D, gen)                      # identical on every rank (same seed)
    p = param_state(X, gen)
    perm = torch.randperm(N_DATA, generator=gen)
    steps_total = args.warmup + args.steps
    global_batch = BATCH * world
    scale = float(N_DATA) / global_batch

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    # the dataset lives in HBM for the device arm; minibatches are gathered by index on the device
    Xd, Yd = X.to(dev), Y.view(-1).to(dev)
    perm_d = perm.to(dev)
    import ctypes as C

    def dev_batch(step):
        lo = ((step * global_batch) + rank * BATCH) % (N_DATA - global_batch)
        idx = perm_d[lo:lo + BATCH]
        return Xd.index_select(0, idx), Yd.index_select(0, idx)

    def measure(compute, sample_clocks):
        """W warm-up + K timed ELBO fwd+bwd steps of the device arm in one compute mode."""
        eng, theta, _, _ = make_engine(p, 'gauss_nonlinear', N_QUAD, dev, compute=compute)
        ei = engine_inputs(p, dev)
        leaves = [ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'], ei['L_raw'], ei['log_var_noise'], theta]
        for t in leaves:
            t.requires_grad_(True)

        def step_fn(xb, yb):
            for t in leaves:
                t.grad = None
            ELL, KLD, _, _, _ = Fn.elbo_terms(eng, xb, yb, scale, ei['Z'], ei['raw_ls'], ei['raw_os'], ei['m'],
                                              ei['L_raw'], ei['log_var_noise'], theta, None, check_status=True)
            loss = -(ELL - KLD)
            loss.backward()
            return loss

        clocks = ClockSampler(local)
        if sample_clocks and rank == 0 and not args.no_clocks:
            clocks.start()                              # sampler runs through warm-up + timed region (same load)
        timing_on = 0 if args.no_gemm_timing else 1
        lib.tgp_gemm_timing(timing_on, None, None)      # events get created during warm-up, not in the timed region
        # pre-warm until the step time is steady (a fresh box starts with cold clocks / lazily loaded modules):
        # stop when three consecutive steps agree within 3 %, or after 8 s
        t_pre, recent = time.perf_counter(), []
        while True:
            t0 = time.perf_counter()
            step_fn(*dev_batch(0))
            torch.cuda.synchronize()
            recent = (recent + [time.perf_counter() - t0])[-3:]
            el = time.perf_counter() - t_pre
            done = el > 8.0 or (len(recent) == 3 and el > 2.0 and max(recent) < 1.03 * min(recent))
            if world > 1:                               # every rank must leave the loop in the same iteration
                flag = torch.tensor([1.0 if done else 0.0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                done = bool(flag.item() > 0.5)
            if done:
                break
        for s in range(args.warmup):
            step_fn(*dev_batch(s))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.tgp_gemm_timing(timing_on, None, None)      # reset the accumulators
        launches0 = lib.tgp_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for s in range(args.warmup, steps_total):
            flush.fill_(s & 0xFF)                       # L2 flush between timed iterations
            loss = step_fn(*dev_batch(s))
        ev1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        gemm_ms, gemm_n = (C.c_double * 3)(), (C.c_long * 3)()
        lib.tgp_gemm_timing(0, gemm_ms, gemm_n)
        launches = lib.tgp_launch_count() - launches0
        clk = clocks.stop() if (sample_clocks and rank == 0) else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # test-NLL forward (no gradients): marginals + quadrature log-lik + moments on the same batches
        with torch.no_grad():
            eng.set_params(*[t.detach() for t in leaves])
            eng.prepare(0.0)
            for s in range(2):
                mu, v = eng.qf_forward(dev_batch(s)[0])
                eng.test_rows(mu, v, dev_batch(s)[1], None, 1, 1.0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(args.steps):
                xb, yb = dev_batch(args.warmup + s)
                eng.prepare(0.0)
                mu, v = eng.qf_forward(xb)
                eng.test_rows(mu, v, yb, None, 1, 1.0)
            e1.record()
            torch.cuda.synchronize()
            e2 = torch.cuda.Event(enable_timing=True)
            for s in range(args.steps):                # same pass with the factorisation reused (frozen parameters)
                xb, yb = dev_batch(args.warmup + s)
                mu, v = eng.qf_forward(xb)
                eng.test_rows(mu, v, yb, None, 1, 1.0)
            e2.record()
            torch.cuda.synchronize()
            tn = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        return dict(ms_total=ms, value=BATCH * world * args.steps / (ms * 1e-3), loss=float(loss.item()),
                    gemm_ms=list(gemm_ms), gemm_n=list(gemm_n), launches=int(launches), clocks=clk,
                    test_nll_rows_per_s=BATCH * world * args.steps / (float(tn[0].item()) * 1e-3),
                    test_nll_cached_rows_per_s=BATCH * world * args.steps / (float(tn[1].item()) * 1e-3))

    # the secondary mode runs first: on a fresh box the first seconds of a process are not steady (cold clocks, lazy
    # module loads), and the headline should not absorb that
    second = 'tf32x3' if args.compute == 'f64' else 'f64'
    measure(second, False)                      # discarded: absorbs the cold start of a fresh box
    head = measure(args.compute, True)
    other = measure(second, False)
    ms_total, value, final_loss, launches, clk = head['ms_total'], head['value'], head['loss'], head['launches'], head['clocks']
    gemm_ms, gemm_n = head['gemm_ms'], head['gemm_n']

    # ---- end-to-end arm: host (pinned) minibatches through the public class API, loss read back every step -----
    e2e = run_e2e(args, p, X, Y, perm, rank, world, dev, scale)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (gemm_f64_kernel, batch contractions) ----------------------------------
    peak64 = measure_fp64_peak(dev)
    alg_flop_step = 6.0 * M * M * BATCH                         # 6*M^2 FLOP per row, fwd+bwd
    gemm_ms_step = (gemm_ms[1] + gemm_ms[2]) / args.steps
    gemm_launches_step = (gemm_n[1] + gemm_n[2]) / args.steps
    achieved = alg_flop_step / (gemm_ms_step * 1e-3) / 1e12 if gemm_ms_step > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    roofline = {'bound': 'tensor', 'kernel': 'gemm_f64_kernel (FP64 DMMA), batch contractions' if args.compute == 'f64'
                else 'gemm_tf32x3_kernel (tcgen05) — event timing covers only the FP64 per-step GEMMs in this mode', 'achieved': achieved,
                'peak': peak64, 'unit': 'TFLOP/s', 'frac': achieved / peak64,
                'peak_source': 'cuBLAS DGEMM 8192^3 via torch.matmul measured in this run (MEASURED_PEAKS.json has no FP64 '
                               'figure; its bf16 %.0f TF/s does not bound an FP64 kernel)' % peaks.get('bf16_tflops', 1707.0),
                'algorithmic_flop_per_launch': alg_flop_step / max(gemm_launches_step, 1),
                'avg_launch_ms': gemm_ms_step / max(gemm_launches_step, 1), 'launches_per_step': gemm_launches_step,
                'kernel_share_of_step': gemm_ms_step / (ms_total / args.steps),
                'executed_dense_tile_tflops': None, 'traffic': None,
                'per_step_o_m3_gemm_ms': gemm_ms[0] / args.steps}
    prof = os.path.join(ROOT, 'profiles', 'r01_roofline_extra.json')
    if os.path.exists(prof):
        try:
            roofline.update(json.load(open(prof)))
        except Exception:
            pass
    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample ----------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        rows_s, sec = cpu_reference_steps(p, X[:CPU_SAMPLE_ROWS * 8], Y[:CPU_SAMPLE_ROWS * 8], CPU_SAMPLE_ROWS, 3, 1)
        cpu = {'value': rows_s, 'unit': 'rows/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '3 steps of a %d-row minibatch of the same workload after 1 warm-up (median %.2f s/step)'
                         % (CPU_SAMPLE_ROWS, sec)}
    line = {'metric': METRIC, 'value': value, 'unit': 'rows/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(world),
            'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
            'final_loss': final_loss, 'compute': args.compute,
            'test_nll': {'value': head['test_nll_rows_per_s'], 'unit': 'rows/s',
                         'what': 'test log-lik + predictive moments forward (prepare + marginals + quadrature), device-resident',
                         'with_factorisation_reused_across_batches': head['test_nll_cached_rows_per_s']},
            'other_mode': {'compute': 'tf32x3' if args.compute == 'f64' else 'f64', 'value': other['value'], 'unit': 'rows/s',
                           'ms_per_step': other['ms_total'] / args.steps, 'final_loss': other['loss'],
                           'loss_rel_diff_vs_headline': abs(other['loss'] - final_loss) / abs(final_loss),
                           'test_nll_rows_per_s': other['test_nll_rows_per_s'],
                           'batch_gemm_ms_per_step': (other['gemm_ms'][1] + other['gemm_ms'][2]) / args.steps,
                           'batch_gemm_algorithmic_tflops': 6.0 * M * M * BATCH / max((other['gemm_ms'][1] + other['gemm_ms'][2]) / args.steps * 1e-3, 1e-12) / 1e12,
                           'note': 'tf32x3 = batch contractions on tcgen05 (3xTF32 split, FP32 TMEM accumulation); per-step '
                                   'factorisation, backward chain and the row epilogue stay FP64 in both modes'}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, p, X, Y, perm, rank, world, dev, scale):
    """Same metric through the public API (`sparse_MF_SP.ELBO` + backward) with HOST buffers: every step gathers its
    minibatch into pinned memory, copies it to the device, and reads the loss back."""
    import torch.distributed as dist
    from tgp.pytorch_b200.dsp import config as cg
    cg.set_maximum_precission()
    cg.device = str(dev)
    from tgp.pytorch_b200.dsp.models import instance_kernel, sparse_MF_SP
    from tgp.pytorch_b200.dsp.models.flow import instance_flow
    from tgp.pytorch_b200.dsp.likelihoods import GaussianNonLinearMean
    from tgp.pytorch_b200.dsp.flows import StepTanhL
    K = instance_kernel('scale_rbf', ard_num_dim=D, num_multioutput=1, kernel_is_shared=False,
                        init_params={'length_scale': 2.0, 'kernel_scale': 2.0})
    lik = GaussianNonLinearMean(out_dim=1, noise_init=0.05, noise_is_shared=False, quadrature_points=N_QUAD)
    np.random.seed(0)
    model = sparse_MF_SP(['zero', K], X[:1024], p['Z'], float(N_DATA), lik, 1, True, False, False, False, False,
                         [instance_flow(StepTanhL(1, 3, add_f0=True))], 'single', 0.0, False,
                         {'variational_distribution': {'variance_scale': 1e-5, 'mean_scale': 0.0}})
    with torch.no_grad():
        model.q_U.variational_mean.copy_(p['m'].view(1, -1))
        model.q_U.chol_variational_covar.copy_(p['L_raw'].view(1, M, M))
        model.covariance_function.base_kernel.raw_lengthscale.copy_(p['raw_lengthscale'].view(1, 1, D))
    model.to(dev)
    model.global_batch_rows = BATCH * world
    from tgp.pytorch_b200.data import PinnedMinibatchStager
    global_batch = BATCH * world
    stager = PinnedMinibatchStager(X, Y, BATCH, dev)

    def batch_index(step):
        lo = ((step * global_batch) + rank * BATCH) % (N_DATA - global_batch)
        return perm[lo:lo + BATCH]

    def one(step):
        """One step of the user-level loop: the step's inputs come from HOST memory (gathered into pinned buffers and
        copied to the device by the stager — for step+1 while step computes), the loss is read back to the host."""
        xb, yb = stager.get()
        model.zero_grad(set_to_none=True)
        ELBO, _, _ = model.ELBO(xb, yb)
        (-ELBO).backward()
        stager.stage(batch_index(step + 1))             # host gather + H2D of the next step under this step's backward
        return float(ELBO.item())                       # device -> host read of the step's result

    stager.stage(batch_index(0))
    for s in range(min(args.warmup, 3)):
        one(s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        one(args.warmup + s)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return {'value': BATCH * world * args.steps / float(dt.item()), 'unit': 'rows/s',
            'h2d_bytes_per_step': stager.bytes_per_step, 'd2h_bytes_per_step': 8,
            'note': 'sparse_MF_SP.ELBO + backward per step through the class API; every step gathers its minibatch from host '
                    'memory into pinned buffers and copies it to the device (PinnedMinibatchStager: the gather + H2D of step '
                    's+1 overlap the kernels of step s), and reads the loss back'}


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1 (e.g. NCCL's version
    banner) was diverted to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--compute', default='f64', choices=['f64', 'tf32x3'], help='f64: DMMA path (what main.py runs); tf32x3: tcgen05 mode')
    ap.add_argument('--no-gemm-timing', action='store_true', help='(diagnostic) do not instrument GEMM launches with events')
    ap.add_argument('--no-clocks', action='store_true', help='(diagnostic) do not sample nvidia-smi during the timed region')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
