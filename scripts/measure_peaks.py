"""cuBLAS peaks that MEASURED_PEAKS.json does not hold: FP64 DGEMM, TF32 and INT8 GEMM at 8192^3 with the
method of that file (best of 10 = burst; back to back for 4 s = sustained).  Writes gpurun_out/peaks_extra.json."""
import json
import os
import time

import torch


def bench(fn, flop):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(10):
            fn()
        n += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return flop / (best * 1e-3) / 1e12, flop * n / (e0.elapsed_time(e1) * 1e-3) / 1e12


def main():
    dev = torch.device('cuda', 0)
    n = 8192
    out = {'gpu_name': torch.cuda.get_device_name(0), 'n': n, 'how': 'torch.matmul / torch._int_mm %d^3, best of 10 (burst) and back to back for 4 s (sustained), CUDA events' % n}
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    out['fp64_tflops'], out['fp64_tflops_sustained'] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    del a, b
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(n, n, dtype=torch.float32, device=dev); b = torch.randn(n, n, dtype=torch.float32, device=dev)
    out['tf32_tflops'], out['tf32_tflops_sustained'] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    torch.backends.cuda.matmul.allow_tf32 = False
    out['fp32_tflops'], out['fp32_tflops_sustained'] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    del a, b
    try:
        a = torch.randint(-64, 64, (n, n), dtype=torch.int8, device=dev); b = torch.randint(-64, 64, (n, n), dtype=torch.int8, device=dev)
        out['int8_tops'], out['int8_tops_sustained'] = bench(lambda: torch._int_mm(a, b), 2.0 * n ** 3)
    except Exception as e:  # noqa: BLE001
        out['int8_tops'] = None
        out['int8_error'] = str(e)[:200]
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/peaks_extra.json', 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
