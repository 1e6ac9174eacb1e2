// Lloyd's k-means on the device for the inducing-point initialisation Z = KMEANS(X_tr, M) (reference code/dsp/utils.py:143-159,
// main.py:145: sklearn KMeans on the host — hours at N = 5 M, M = 1024).  SURVEY.md §8f rank 4.
//   k_kmeans_assign: a thread per row keeps its point in registers and walks the centroids through shared-memory tiles;
//                    writes the nearest centroid, accumulates the inertia, and adds the point to per-CTA partial sums of the
//                    centroids it hit (shared-memory atomics when M * D fits, global atomics otherwise);
//   k_kmeans_update: centroid = sum / count; an empty cluster keeps its previous centre.
#pragma once
#include "common.cuh"

namespace tgp {

constexpr int KM_THREADS = 256, KM_TILE = 128;

template <int MAXD>
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_assign(const double* __restrict__ X, long N, int D, const double* __restrict__ C,
                                                              int M, int* __restrict__ assign, double* __restrict__ sums,
                                                              double* __restrict__ counts, double* __restrict__ inertia, int use_smem) {
    extern __shared__ double sm_km[];
    double* ctile = sm_km;                                   // [KM_TILE][MAXD]
    double* psum = sm_km + KM_TILE * MAXD;                   // [M][D + 1] partial sums + counts of this CTA (use_smem)
    if (use_smem) for (int i = threadIdx.x; i < M * (D + 1); i += KM_THREADS) psum[i] = 0.0;
    const long n = (long)blockIdx.x * KM_THREADS + threadIdx.x;
    double x[MAXD];
#pragma unroll
    for (int d = 0; d < MAXD; ++d) x[d] = (n < N && d < D) ? X[n * D + d] : 0.0;
    double best = INFINITY;
    int arg = 0;
    for (int c0 = 0; c0 < M; c0 += KM_TILE) {
        __syncthreads();
        for (int i = threadIdx.x; i < KM_TILE * MAXD; i += KM_THREADS) {
            const int c = i / MAXD, d = i % MAXD;
            ctile[i] = (c0 + c < M && d < D) ? C[(long)(c0 + c) * D + d] : 0.0;
        }
        __syncthreads();
        const int lim = min(KM_TILE, M - c0);
        for (int c = 0; c < lim; ++c) {
            double dist = 0.0;
#pragma unroll
            for (int d = 0; d < MAXD; ++d) { const double df = x[d] - ctile[c * MAXD + d]; dist = fma(df, df, dist); }
            if (dist < best) { best = dist; arg = c0 + c; }
        }
    }
    double local = 0.0;
    if (n < N) {
        if (assign) assign[n] = arg;
        local = best;
        if (use_smem) {
            for (int d = 0; d < D; ++d) atomicAdd(&psum[arg * (D + 1) + d], x[d]);
            atomicAdd(&psum[arg * (D + 1) + D], 1.0);
        } else {
            for (int d = 0; d < D; ++d) atomicAdd(&sums[(long)arg * D + d], x[d]);
            atomicAdd(&counts[arg], 1.0);
        }
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0 && local != 0.0) atomicAdd(inertia, local);
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < M * (D + 1); i += KM_THREADS) {
            const double v = psum[i];
            if (v != 0.0) {
                const int c = i / (D + 1), d = i % (D + 1);
                if (d < D) atomicAdd(&sums[(long)c * D + d], v); else atomicAdd(&counts[c], v);
            }
        }
    }
}

__global__ void k_kmeans_update(double* __restrict__ C, const double* __restrict__ sums, const double* __restrict__ counts, int M, int D) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M * D) {
        const double cnt = counts[i / D];
        if (cnt > 0.0) C[i] = sums[i] / cnt;
    }
}

// one Lloyd iteration: sums / counts / inertia are zeroed here; centroids are updated in place unless update == 0
inline int kmeans_iteration(const double* X, long N, int D, double* C, int M, int* assign, double* sums, double* counts,
                            double* inertia, int update, cudaStream_t st) {
    if (D > 32) return set_error(-2, "device k-means supports input dimension <= 32");
    cudaMemsetAsync(sums, 0, (size_t)M * D * sizeof(double), st);
    cudaMemsetAsync(counts, 0, (size_t)M * sizeof(double), st);
    cudaMemsetAsync(inertia, 0, sizeof(double), st);
    const int maxd = D <= 4 ? 4 : (D <= 8 ? 8 : (D <= 16 ? 16 : 32));
    const size_t tile = (size_t)KM_TILE * maxd * sizeof(double), part = (size_t)M * (D + 1) * sizeof(double);
    const int use_smem = tile + part <= 160 * 1024 ? 1 : 0;
    const size_t smem = tile + (use_smem ? part : 0);
    const unsigned grid = (unsigned)cdiv(N, KM_THREADS);
#define TGP_KM(MD) do { \
        cudaFuncSetAttribute(k_kmeans_assign<MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
        k_kmeans_assign<MD><<<grid, KM_THREADS, smem, st>>>(X, N, D, C, M, assign, sums, counts, inertia, use_smem); } while (0)
    if (maxd == 4) TGP_KM(4); else if (maxd == 8) TGP_KM(8); else if (maxd == 16) TGP_KM(16); else TGP_KM(32);
#undef TGP_KM
    TGP_TRY(check_launch("k_kmeans_assign"));
    if (update) {
        k_kmeans_update<<<(unsigned)cdiv((long)M * D, 256), 256, 0, st>>>(C, sums, counts, M, D);
        TGP_TRY(check_launch("k_kmeans_update"));
    }
    return 0;
}

}  // namespace tgp
