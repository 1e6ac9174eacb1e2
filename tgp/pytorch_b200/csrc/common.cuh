// Shared helpers for the tgp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define TGP_WARP 32

namespace tgp {

// last error text, returned by tgp_last_error()
extern char g_last_error[512];

inline int set_error(int code, const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
    return code;
}

extern long g_launch_count;      // kernels launched by this library (tgp_launch_count())

inline int check_launch(const char* what) {
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
        return -100;
    }
    return 0;
}

#define TGP_TRY(expr) do { int _rc = (expr); if (_rc != 0) return _rc; } while (0)

// A library-owned side stream per device, used to run K_xz generation concurrently with the (single-SM-bound) factorisation:
// tgp_prepare records `params_ready` right after the parameter transforms; tgp_qf_forward generates K_xz on the side stream
// as soon as that event fires and joins back with `k_ready`.  Host objects only (no device memory).  `capturing` remembers
// whether the prepare was enqueued under CUDA-graph capture: fork and join must belong to the same capture.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t params_ready = nullptr, k_ready = nullptr;
    bool have_params = false, capturing = false;
    bool fresh = false;          // set by tgp_prepare, consumed by the FIRST forward after it: a later forward on the same
                                 // factorisation may follow work that still reads the batch workspace, and stays on the main stream
};
inline SideStream& side_stream() {
    static SideStream s[64];
    int dev = 0;
    cudaGetDevice(&dev);
    SideStream& x = s[dev >= 0 && dev < 64 ? dev : 0];
    if (!x.stream) {
        int lo = 0, hi = 0;          // lowest priority: the factorisation's small kernels are never queued behind K_xz tiles
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&x.stream, cudaStreamNonBlocking, lo);
        cudaEventCreateWithFlags(&x.params_ready, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&x.k_ready, cudaEventDisableTiming);
    }
    return x;
}
inline bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs != cudaStreamCaptureStatusNone;
}
extern int g_overlap_kgen;       // TGP_OPT_OVERLAP_KGEN

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__host__ __device__ inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per DEVICE: one-time kernel configuration is tracked per (site, device), so that a second GPU
// used from the same process gets its own opt-in to large dynamic shared memory.
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

// softplus and its derivative as torch.nn.functional.softplus (beta=1, threshold=20) evaluates them
__device__ __forceinline__ double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }
__device__ __forceinline__ double sigmoid_d(double x) { return x > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-x)); }

}  // namespace tgp
