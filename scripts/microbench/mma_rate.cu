// Issue-rate microbenchmark of tcgen05.mma on sm_100a: kind::i8 (s8 x s8 -> s32) against kind::tf32 / kind::f16, for
// the accumulator patterns the sliced-integer FP64 contraction (csrc/gemm_i8.cuh) uses.  Shared-memory contents are
// arbitrary (only the rate matters).  One CTA per SM, one issuing thread, K-major 128B-swizzled operand descriptors.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(const void* smem_tile) {
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if (KIND == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    else if (KIND == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// KIND 0: i8 (K = 32 per MMA), 1: tf32 (K = 8), 2: bf16 (K = 16); N columns per MMA; NACC accumulators used round-robin;
// NSLICE distinct A / B operand tiles per stage (slice pairs p+q = antidiagonal pattern when PATTERN = 1)
template <int KIND, int N, int NACC, int PATTERN>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* cycles, int nslice) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const int A_BYTES = 128 * 128, B_BYTES = N * 128;          // 128-byte rows (one swizzle atom along K)
    uint32_t idesc;
    if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    else if (KIND == 1) idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    else idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 32) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            // one "stage": 4 k-steps of 32 bytes inside the swizzle row
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                if (PATTERN == 1) {
                    for (int s = 0; s < nslice; ++s)
                        for (int p = 0; p <= s; ++p) {
                            const uint64_t dA = make_desc(smem + p * A_BYTES) + adv;
                            const uint64_t dB = make_desc(smem + nslice * A_BYTES + (s - p) * B_BYTES) + adv;
                            umma<KIND>(tmem_base + (uint32_t)((s % NACC) * N), dA, dB, idesc, (it | kk | p) ? 1u : 0u);
                        }
                } else {
                    for (int s = 0; s < nslice; ++s) {
                        const uint64_t dA = make_desc(smem + (s % 2) * A_BYTES) + adv;
                        const uint64_t dB = make_desc(smem + 2 * A_BYTES + (s % 2) * B_BYTES) + adv;
                        umma<KIND>(tmem_base + (uint32_t)((s % NACC) * N), dA, dB, idesc, (it | kk) ? 1u : 0u);
                    }
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n"
                     ::"r"(smem_u32(&bar)), "r"(0) : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

template <int KIND, int N, int NACC, int PATTERN>
void run(const char* name, int nslice, int kelems) {
    const int smem = 1024 + nslice * (128 * 128 + N * 128) + 4096;
    cudaFuncSetAttribute(k_rate<KIND, N, NACC, PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* cyc;
    cudaMalloc(&cyc, 8);
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rate<KIND, N, NACC, PATTERN><<<148, 128, smem>>>(10, cyc, nslice);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_rate<KIND, N, NACC, PATTERN><<<148, 128, smem>>>(iters, cyc, nslice);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const long mmas = (long)iters * 4 * (PATTERN == 1 ? nslice * (nslice + 1) / 2 : nslice);
    const double ops = 2.0 * 128 * N * kelems * (double)mmas * 148;
    printf("%-44s %s  %8.3f ms  %8.1f Tops/s  %6.1f clk/MMA (SM0)\n", name, err == cudaSuccess ? "ok " : cudaGetErrorString(err), ms,
           ops / (ms * 1e-3) / 1e12, (double)c / (double)mmas);
    cudaFree(cyc);
}

int main() {
    run<0, 256, 2, 0>("i8   N=256 2 acc, plain", 4, 32);
    run<0, 128, 4, 0>("i8   N=128 4 acc, plain", 4, 32);
    run<0, 64, 8, 0>("i8   N=64  8 acc, plain", 8, 32);
    run<0, 64, 8, 1>("i8   N=64  8 acc, 8 slices antidiag (36 MMA)", 8, 32);
    run<0, 64, 8, 1>("i8   N=64  7 slices antidiag (28 MMA)", 7, 32);
    run<0, 128, 4, 1>("i8   N=128 4 slices antidiag (10 MMA)", 4, 32);
    run<0, 32, 8, 1>("i8   N=32  8 slices antidiag", 8, 32);
    run<1, 256, 2, 0>("tf32 N=256 2 acc, plain", 4, 8);
    run<2, 256, 2, 0>("bf16 N=256 2 acc, plain", 4, 16);
    return 0;
}
