// FP64 tile GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64) for sm_100a.
//
//   C[m,n] = alpha * sum_k Aop[m,k] * Bop[n,k] + beta * C[m,n]          (both operands indexed [row, k])
//
// One CTA = 128x128 output tile, 256 threads (8 warps as 2x4, warp tile 64x32 = 8x4 DMMA tiles), BK = 16,
// register-staged double buffering through two shared-memory stages.  Operands may be stored k-contiguous
// ([row][k]) or row-contiguous ([k][row]); shared memory always holds [row][k] with a +4 pad, which makes the
// 8x4 / 4x8 DMMA fragment reads (one LDS.64 per lane) bank-conflict free.
//
// Triangular structure is exploited by clipping the k-range per tile (a_tri / b_tri) and by skipping tiles
// strictly above the diagonal (c_lower).  The masked triangle of a "triangular" operand must hold zeros.
//
// This one kernel carries every FP64 contraction of the path: the blocked Cholesky panel/SYRK updates, the
// recursive-doubling triangular inverse, C = L_S^T L^-1, the batch contractions A = K L^-T, B = K C^T, their
// data- and weight-gradient mirrors, and the Cholesky-backward chain.
#pragma once
#include "common.cuh"

namespace tgp {

struct GemmArgs {
    int M, N, K;
    const double* A; long lda; long strideA;
    const double* B; long ldb; long strideB;
    double* C; long ldc; long strideC;
    double alpha, beta;
    int a_layout;   // 0: A[m*lda + k]   1: A[k*lda + m]
    int b_layout;   // 0: B[n*ldb + k]   1: B[k*ldb + n]
    int a_tri;      // 0 dense, 1: Aop[m,k] != 0 only for k <= m, 2: only for k >= m
    int b_tri;      // 0 dense, 1: Bop[n,k] != 0 only for k <= n, 2: only for k >= n
    int c_lower;    // 1: write only elements with n <= m (tiles strictly above the diagonal are skipped)
    int batch;
    int splitk;     // > 1: blockIdx.z splits the k-range; partial sums are atomically ADDED to C (beta must be 1, batch 1)
    int tag;        // profiling class: 0 = per-step O(M^3) work, 1 = batch contraction (rows x M x M)
};

// Optional in-stream timing of every GEMM launch (bench.py's live roofline measurement): a pair of CUDA events per
// launch on the launching stream, resolved at query time.  Off by default.
struct GemmTimer {
    bool enabled = false;
    static constexpr int MAXEV = 8192;
    cudaEvent_t ev[2 * MAXEV];
    int tagv[MAXEV];
    int n = 0, created = 0;
    double ms[3] = {0, 0, 0};
    long launches[3] = {0, 0, 0};
    void begin(int tag, cudaStream_t st) {
        if (n >= MAXEV) flush();
        while (created < 2 * (n + 1)) { cudaEventCreate(&ev[created]); ++created; }
        tagv[n] = tag;
        cudaEventRecord(ev[2 * n], st);
    }
    void end(cudaStream_t st) { cudaEventRecord(ev[2 * n + 1], st); ++n; }
    void flush() {
        for (int i = 0; i < n; ++i) {
            float t = 0.f;
            cudaEventSynchronize(ev[2 * i + 1]);
            cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
            ms[tagv[i]] += t; ++launches[tagv[i]];
        }
        n = 0;
    }
};
extern GemmTimer g_gemm_timer;

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4, GLD = GBK + GPAD;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM_BYTES = 2 * 2 * GBM * GLD * (int)sizeof(double);   // 2 stages x (A,B)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Loads this thread's 8 elements of a 128 x 16 operand tile into r[8].
//   layout 0: thread -> row = tid>>1, k-offset (tid&1)*8, 8 consecutive k
//   layout 1: thread -> k = tid&15, rows (tid>>4)*8 .. +7   (k fastest across lanes: transposing smem stores
//             then hit 16 distinct banks)
template <int LAYOUT>
__device__ __forceinline__ void load_tile(double (&r)[8], const double* __restrict__ P, long ld, int rows, int K,
                                          int row0, int k0, int kend, bool vec_ok, int tid) {
    if (LAYOUT == 0) {
        const int row = row0 + (tid >> 1);
        const int k = k0 + (tid & 1) * 8;
        if (row < rows && vec_ok && k + 8 <= kend) {
            const double2* src = reinterpret_cast<const double2*>(P + (long)row * ld + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) { double2 t = __ldg(src + i); r[2 * i] = t.x; r[2 * i + 1] = t.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = (row < rows && k + i < kend) ? __ldg(P + (long)row * ld + k + i) : 0.0;
        }
    } else {
        const int k = k0 + (tid & 15);
        const int row = row0 + (tid >> 4) * 8;
        if (k < kend && vec_ok && row + 8 <= rows) {
            const double2* src = reinterpret_cast<const double2*>(P + (long)k * ld + row);
#pragma unroll
            for (int i = 0; i < 4; ++i) { double2 t = __ldg(src + i); r[2 * i] = t.x; r[2 * i + 1] = t.y; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = (k < kend && row + i < rows) ? __ldg(P + (long)k * ld + row + i) : 0.0;
        }
    }
}

template <int LAYOUT>
__device__ __forceinline__ void store_tile(const double (&r)[8], double* __restrict__ S, int tid) {
    if (LAYOUT == 0) {
        double2* dst = reinterpret_cast<double2*>(S + (tid >> 1) * GLD + (tid & 1) * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_double2(r[2 * i], r[2 * i + 1]);
    } else {
        const int k = tid & 15, row = (tid >> 4) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) S[(row + i) * GLD + k] = r[i];
    }
}

template <int AL, int BL>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_f64_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    // m-tiles vary fastest (CTAs that are co-resident share the B tile); n-tiles are visited heaviest-first when the
    // triangular clipping makes their k-extent grow with n (b_tri == 1), so that the last wave holds the light tiles
    const int nt = (g.b_tri == 1) ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
    const int m0 = blockIdx.x * GBM, n0 = nt * GBN;
    if (g.c_lower && n0 > m0 + GBM - 1) return;

    const int zb = g.splitk > 1 ? 0 : blockIdx.z;
    const double* A = g.A + (long)zb * g.strideA;
    const double* B = g.B + (long)zb * g.strideB;
    double* C = g.C + (long)zb * g.strideC;

    int kb = 0, ke = g.K;
    if (g.a_tri == 1) ke = min(ke, m0 + GBM);
    if (g.a_tri == 2) kb = max(kb, m0);
    if (g.b_tri == 1) ke = min(ke, n0 + GBN);
    if (g.b_tri == 2) kb = max(kb, n0);
    kb = (kb / GBK) * GBK;
    int nk = ke > kb ? (ke - kb + GBK - 1) / GBK : 0;
    if (g.splitk > 1) {            // this CTA's share of the k-tiles
        const int per = (nk + g.splitk - 1) / g.splitk;
        const int t0 = min(nk, (int)blockIdx.z * per), t1 = min(nk, t0 + per);
        kb += t0 * GBK;
        ke = min(ke, kb + (t1 - t0) * GBK);
        nk = t1 - t0;
        if (nk == 0) return;
    }

    const bool vecA = ((g.lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool vecB = ((g.ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // stage s: A tile at smem + s*STAGE, B tile at smem + s*STAGE + GBM*GLD
    constexpr int STAGE = 2 * GBM * GLD;
    double ra[8], rb[8];

    if (nk > 0) {
        load_tile<AL>(ra, A, g.lda, g.M, g.K, m0, kb, ke, vecA, tid);
        load_tile<BL>(rb, B, g.ldb, g.N, g.K, n0, kb, ke, vecB, tid);
        store_tile<AL>(ra, smem, tid);
        store_tile<BL>(rb, smem + GBM * GLD, tid);
    }
    __syncthreads();

    const int fr = lane >> 2, fc = lane & 3;
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile<AL>(ra, A, g.lda, g.M, g.K, m0, kb + (kt + 1) * GBK, ke, vecA, tid);
            load_tile<BL>(rb, B, g.ldb, g.N, g.K, n0, kb + (kt + 1) * GBK, ke, vecB, tid);
        }
        const double* as = smem + cur * STAGE + (wm * 64 + fr) * GLD + fc;
        const double* bs = smem + cur * STAGE + GBM * GLD + (wn * 32 + fr) * GLD + fc;
#pragma unroll
        for (int k4 = 0; k4 < GBK / 4; ++k4) {
            double fa[8], fb[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) fa[i] = as[i * 8 * GLD + k4 * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) fb[j] = bs[j * 8 * GLD + k4 * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
        if (kt + 1 < nk) {
            store_tile<AL>(ra, smem + (cur ^ 1) * STAGE, tid);
            store_tile<BL>(rb, smem + (cur ^ 1) * STAGE + GBM * GLD, tid);
        }
        __syncthreads();
    }

    // epilogue: each lane owns C[row = fr][cols 2*fc, 2*fc+1] of every 8x8 tile
    const bool vecC = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + wm * 64 + i * 8 + fr;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn * 32 + j * 8 + 2 * fc;
            if (n >= g.N) continue;
            double* cp = C + (long)m * g.ldc + n;
            double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
            const bool ok0 = !g.c_lower || n <= m;
            const bool ok1 = (n + 1 < g.N) && (!g.c_lower || n + 1 <= m);
            if (g.splitk > 1) {
                if (ok0) atomicAdd(cp, v0);
                if (ok1) atomicAdd(cp + 1, v1);
            } else if (ok0 && ok1 && vecC) {
                if (g.beta != 0.0) { double2 o = *reinterpret_cast<double2*>(cp); v0 += g.beta * o.x; v1 += g.beta * o.y; }
                *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
            } else {
                if (ok0) { if (g.beta != 0.0) v0 += g.beta * cp[0]; cp[0] = v0; }
                if (ok1) { if (g.beta != 0.0) v1 += g.beta * cp[1]; cp[1] = v1; }
            }
        }
    }
}

inline int gemm_f64(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    if (g.splitk > 1 && (g.batch != 1 || g.beta != 1.0)) return set_error(-3, "split-k GEMM needs batch == 1 and beta == 1");
    dim3 grid((unsigned)cdiv(g.M, GBM), (unsigned)cdiv(g.N, GBN), (unsigned)(g.splitk > 1 ? g.splitk : g.batch));
    if (grid.y > 65535) return set_error(-3, "too many column tiles");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_f64_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        cudaFuncSetAttribute(gemm_f64_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        cudaFuncSetAttribute(gemm_f64_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        cudaFuncSetAttribute(gemm_f64_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        attr_set = true;
    }
    const bool timed = g_gemm_timer.enabled;
    if (timed) g_gemm_timer.begin(g.tag, st);
    if (g.a_layout == 0 && g.b_layout == 0) gemm_f64_kernel<0, 0><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(g);
    else if (g.a_layout == 0 && g.b_layout == 1) gemm_f64_kernel<0, 1><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(g);
    else if (g.a_layout == 1 && g.b_layout == 0) gemm_f64_kernel<1, 0><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(g);
    else gemm_f64_kernel<1, 1><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(g);
    if (timed) g_gemm_timer.end(st);
    return check_launch("gemm_f64");
}

// convenience builder
inline GemmArgs make_gemm(int M, int N, int K, const double* A, long lda, int al, const double* B, long ldb, int bl,
                          double* C, long ldc, double alpha = 1.0, double beta = 0.0) {
    GemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.strideA = 0;
    g.B = B; g.ldb = ldb; g.strideB = 0;
    g.C = C; g.ldc = ldc; g.strideC = 0;
    g.alpha = alpha; g.beta = beta;
    g.a_layout = al; g.b_layout = bl;
    g.a_tri = 0; g.b_tri = 0; g.c_lower = 0; g.batch = 1; g.splitk = 1; g.tag = 0;
    return g;
}

}  // namespace tgp
