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

// CTA tile 128 x 64 (4 warps as 2 x 2, warp tile 64 x 32), two CTAs per SM: one CTA's prologue / epilogue / barrier bubbles
// are covered by the other's MMAs (the 128 accumulator registers per thread leave room for exactly 8 such warps per SM).
// -DTGP_GEMM_WIDE selects the earlier 128 x 128 / 8-warp / one-CTA-per-SM shape for A/B measurements.
#ifdef TGP_GEMM_WIDE
constexpr int GBM = 128, GBN = 128, GBK = 16, GSTAGES = 4, GEMM_CTAS_PER_SM = 1;
#else
constexpr int GBM = 128, GBN = 64, GBK = 16, GSTAGES = 3, GEMM_CTAS_PER_SM = 2;
#endif
constexpr int GEMM_GROUP_M = 32;     // m-tiles per rasterisation group
constexpr int GWARPS_N = GBN / 32;
constexpr int GEMM_THREADS = 32 * 2 * GWARPS_N;
constexpr int GEMM_SLOTS = 148 * GEMM_CTAS_PER_SM;      // co-resident CTAs on a B200
constexpr int GLD = GBK + 4;        // [row][k] tile: row pitch in doubles (conflict-free 8x4 fragment reads)
constexpr int GLDT_A = GBM + 4;     // [k][row] tiles: k pitch in doubles, = 4 mod 16 (conflict-free 4x8 fragment reads)
constexpr int GLDT_B = GBN + 4;
constexpr int GOPER_A = GBM * GLD > GBK * GLDT_A ? GBM * GLD : GBK * GLDT_A;    // doubles per operand slot
constexpr int GOPER_B = GBN * GLD > GBK * GLDT_B ? GBN * GLD : GBK * GLDT_B;
constexpr int GSTAGE = GOPER_A + GOPER_B;
constexpr int GEMM_SMEM_BYTES = GSTAGES * GSTAGE * (int)sizeof(double);

// number of CTAs that do work for an M x N output (lower: tiles strictly above the diagonal exit immediately)
inline long gemm_tiles(int M, int N, bool lower) {
    const long Tm = (M + GBM - 1) / GBM, Tn = (N + GBN - 1) / GBN;
    if (!lower) return Tm * Tn;
    long t = 0;
    for (long i = 0; i < Tm; ++i)
        for (long j = 0; j < Tn; ++j) t += (j * GBN <= i * GBM + GBM - 1);
    return t;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// cp.async with zero fill: copies `bytes` (0..16 / 0..8) from global and zero-fills the rest of the destination.
__device__ __forceinline__ void cp_async16(double* dst, const double* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_async8(double* dst, const double* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Asynchronously copies one ROWS x 16 operand tile (rows row0.., k-range k0..k0+15 clipped to kend) into shared memory.
//   layout 0 (memory [row][k]):  smem [row][GLD];      16-byte chunk c -> row c>>3, k (c&7)*2
//   layout 1 (memory [k][row]):  smem [k][ROWS + 4];   16-byte chunk c -> k c/(ROWS/2), row (c%(ROWS/2))*2
// Out-of-range elements are zero-filled by the copy itself; unaligned operands fall back to 8-byte copies.
template <int LAYOUT, int ROWS>
__device__ __forceinline__ void issue_tile(double* __restrict__ S, const double* __restrict__ P, long ld, int rows,
                                           int row0, int k0, int kend, bool vec_ok, int tid) {
    constexpr int PITCH_T = ROWS + 4;
    static_assert((ROWS * 8) % GEMM_THREADS == 0, "tile chunks must divide over the threads");
#pragma unroll
    for (int i = 0; i < ROWS * 8 / GEMM_THREADS; ++i) {
        const int c = tid + i * GEMM_THREADS;
        int r, k;
        double* dst;
        if (LAYOUT == 0) { r = c >> 3; k = (c & 7) * 2; dst = S + r * GLD + k; }
        else { k = c / (ROWS / 2); r = (c % (ROWS / 2)) * 2; dst = S + k * PITCH_T + r; }
        const int row = row0 + r, kk = k0 + k;
        // number of valid elements along the contiguous direction (0, 1 or 2); 0 if the other index is out of range
        int nv;
        const double* src;
        if (LAYOUT == 0) { nv = row < rows ? min(2, max(0, kend - kk)) : 0; src = P + (long)row * ld + kk; }
        else { nv = kk < kend ? min(2, max(0, rows - row)) : 0; src = P + (long)kk * ld + row; }
        if (nv == 0) src = P;
        if (vec_ok) {
            cp_async16(dst, src, nv * 8);
        } else {
            cp_async8(dst, src, nv >= 1 ? 8 : 0);
            cp_async8(dst + 1, nv >= 2 ? src + 1 : P, nv >= 2 ? 8 : 0);
        }
    }
}

// Fast path of the same copy for tiles that are interior in rows and k (every tile of the batch contractions at the
// benchmark sizes): all address arithmetic is done once per CTA, a k-tile costs one pointer bump per operand and one
// cp.async per 16-byte chunk.  (The generic path above spends ~40 integer instructions per chunk on bounds.)
template <int LAYOUT, int ROWS>
struct TileCopy {
    static constexpr int NCHUNK = ROWS * 8 / GEMM_THREADS;
    static constexpr int DSTEP = LAYOUT == 0 ? (GEMM_THREADS / 8) * GLD : (GEMM_THREADS / (ROWS / 2)) * (ROWS + 4);
    const double* src0;   // this thread's chunk 0 of k-tile 0
    long kstep, istep;    // doubles per k-tile / per chunk index
    int dst0;
    __device__ __forceinline__ void init(const double* P, long ld, int row0, int kb, int tid) {
        if (LAYOUT == 0) {
            const int r = tid >> 3, k = (tid & 7) * 2;
            src0 = P + (long)(row0 + r) * ld + kb + k; kstep = GBK; istep = (long)(GEMM_THREADS / 8) * ld; dst0 = r * GLD + k;
        } else {
            const int k = tid / (ROWS / 2), r = (tid % (ROWS / 2)) * 2;
            src0 = P + (long)(kb + k) * ld + row0 + r; kstep = (long)GBK * ld; istep = (long)(GEMM_THREADS / (ROWS / 2)) * ld;
            dst0 = k * (ROWS + 4) + r;
        }
    }
    __device__ __forceinline__ void issue(double* S, int kt) const {
        const double* src = src0 + (long)kt * kstep;
        double* dst = S + dst0;
#pragma unroll
        for (int i = 0; i < NCHUNK; ++i) cp_async16(dst + i * DSTEP, src + (long)i * istep, 16);
    }
};

template <int AL, int BL>
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_CTAS_PER_SM) gemm_f64_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Rasterisation: CTAs are dealt in groups of GEMM_GROUP_M m-tiles x all n-tiles, m fastest inside a group (co-resident
    // CTAs share the B tile; the A rows of a group, <= 4096 x K doubles, are read from HBM once and served to the other
    // n-tiles by L2).  n-tiles are visited heaviest-first when the triangular clipping makes their k-extent grow with n
    // (b_tri == 1), so that the last wave holds the light tiles.
    int mt, nt;
    {
        const int Tm = (int)gridDim.x, Tn = (int)gridDim.y;
        const int lin = (int)blockIdx.y * Tm + (int)blockIdx.x;
        const int grp = lin / (GEMM_GROUP_M * Tn), within = lin - grp * GEMM_GROUP_M * Tn;
        const int g_rows = min(GEMM_GROUP_M, Tm - grp * GEMM_GROUP_M);
        nt = within / g_rows;
        mt = grp * GEMM_GROUP_M + (within - nt * g_rows);
    }
    if (g.b_tri == 1) nt = (int)gridDim.y - 1 - nt;
    const int m0 = mt * GBM, n0 = nt * GBN;
    // warp grid 2 (m) x GWARPS_N (n).  Warps that share a scheduler get complementary n-positions, so that the per-warp
    // triangular clipping below removes the same amount of work from every scheduler: in the 8-warp shape warp and
    // warp + 4 take n-quarters q and 3 - q; in the 4-warp shape neighbouring m-tiles (co-resident CTAs) swap halves.
#ifdef TGP_GEMM_WIDE
    const int wm = warp >> 2, wn = (warp < 4) ? warp : 7 - warp;
#else
    const int wm = warp >> 1, wn = (warp & 1) ^ (mt & 1);
#endif
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
    // per-warp k-range inside the tile's range: a warp whose 64 rows / 32 columns see only zeros of a triangular operand
    // in a k-tile skips that k-tile's MMAs (it still takes part in the copies and barriers)
    int wkb = kb, wke = ke;
    if (g.a_tri == 1) wke = min(wke, m0 + wm * 64 + 64);
    if (g.a_tri == 2) wkb = max(wkb, m0 + wm * 64);
    if (g.b_tri == 1) wke = min(wke, n0 + wn * 32 + 32);
    if (g.b_tri == 2) wkb = max(wkb, n0 + wn * 32);
    const bool warp_dead = g.c_lower && (n0 + wn * 32 > m0 + wm * 64 + 63);     // sub-tile strictly above the diagonal

    const bool vecA = ((g.lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool vecB = ((g.ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // beta * C enters through the accumulators: the read of C is issued before the k-loop and hidden behind it, instead
    // of a dependent read-modify-write at the very end of the CTA
    const int fr = lane >> 2, fc = lane & 3;
    const bool vecC = ((g.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const bool beta_early = g.beta != 0.0 && g.splitk <= 1 && g.alpha != 0.0 && !warp_dead;
    if (beta_early) {
        const double sc = g.beta / g.alpha;
        if (vecC && m0 + GBM <= g.M && n0 + GBN <= g.N) {
            // interior tile: 32 independent, unconditional 16-byte loads in flight at once (elements above the diagonal of
            // a lower-only output are read and masked — they exist in memory, they are just never written)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double2 o = *reinterpret_cast<const double2*>(C + (long)(m0 + wm * 64 + i * 8 + fr) * g.ldc + n0 + wn * 32 + j * 8 + 2 * fc);
                    acc[i][j][0] = o.x; acc[i][j][1] = o.y;
                }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int m = m0 + wm * 64 + i * 8 + fr, n = n0 + wn * 32 + j * 8 + 2 * fc;
                    acc[i][j][0] = (!g.c_lower || n <= m) ? sc * acc[i][j][0] : 0.0;
                    acc[i][j][1] = (!g.c_lower || n + 1 <= m) ? sc * acc[i][j][1] : 0.0;
                }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + wm * 64 + i * 8 + fr;
                if (m >= g.M) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = n0 + wn * 32 + j * 8 + 2 * fc;
                    if (n >= g.N) continue;
                    const double* cp = C + (long)m * g.ldc + n;
                    if (!g.c_lower || n <= m) acc[i][j][0] = sc * cp[0];
                    if ((n + 1 < g.N) && (!g.c_lower || n + 1 <= m)) acc[i][j][1] = sc * cp[1];
                }
            }
        }
    }

    TileCopy<AL, GBM> cpA;
    TileCopy<BL, GBN> cpB;
    cpA.init(A, g.lda, m0, kb, tid);
    cpB.init(B, g.ldb, n0, kb, tid);
    const bool fastA = vecA && m0 + GBM <= g.M, fastB = vecB && n0 + GBN <= g.N;
    auto load_ktile = [&](int t) {                 // k-tile t -> its ring slot
        double* S = smem + (t % GSTAGES) * GSTAGE;
        const bool k_inner = kb + (t + 1) * GBK <= ke;
        if (fastA && k_inner) cpA.issue(S, t);
        else issue_tile<AL, GBM>(S, A, g.lda, g.M, m0, kb + t * GBK, ke, vecA, tid);
        if (fastB && k_inner) cpB.issue(S + GOPER_A, t);
        else issue_tile<BL, GBN>(S + GOPER_A, B, g.ldb, g.N, n0, kb + t * GBK, ke, vecB, tid);
    };
    // prologue: GSTAGES-1 k-tiles in flight (one commit group per k-tile, empty groups keep the counting uniform)
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < nk) load_ktile(s);
        cp_async_commit();
    }

    // fragment element (row r, k) of a tile: layout 0 -> r*GLD + k, layout 1 -> k*pitch + r
    const int a_off = (AL == 0) ? (wm * 64 + fr) * GLD + fc : fc * GLDT_A + wm * 64 + fr;
    const int b_off = (BL == 0) ? (wn * 32 + fr) * GLD + fc : fc * GLDT_B + wn * 32 + fr;
    constexpr int A_RS = (AL == 0) ? 8 * GLD : 8, A_KS = (AL == 0) ? 4 : 4 * GLDT_A;
    constexpr int B_RS = (BL == 0) ? 8 * GLD : 8, B_KS = (BL == 0) ? 4 : 4 * GLDT_B;

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GSTAGES - 2>();          // k-tile kt has landed (for this thread's copies)
        __syncthreads();                       // ... and for everyone's; everyone is also done with k-tile kt-1
        {
            const int nx = kt + GSTAGES - 1;   // refill the slot k-tile kt-1 occupied
            if (nx < nk) load_ktile(nx);
            cp_async_commit();
        }
        const int k0 = kb + kt * GBK;
        if (warp_dead || k0 >= wke || k0 + GBK <= wkb) continue;
        const double* as = smem + (kt % GSTAGES) * GSTAGE + a_off;
        const double* bs = smem + (kt % GSTAGES) * GSTAGE + GOPER_A + b_off;
#pragma unroll
        for (int k4 = 0; k4 < GBK / 4; ++k4) {
            double fa[8], fb[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) fa[i] = as[i * A_RS + k4 * A_KS];
#pragma unroll
            for (int j = 0; j < 4; ++j) fb[j] = bs[j * B_RS + k4 * B_KS];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each lane owns C[row = fr][cols 2*fc, 2*fc+1] of every 8x8 tile
    if (warp_dead) return;
    const double beta = beta_early ? 0.0 : g.beta;
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
                if (beta != 0.0) { double2 o = *reinterpret_cast<double2*>(cp); v0 += beta * o.x; v1 += beta * o.y; }
                *reinterpret_cast<double2*>(cp) = make_double2(v0, v1);
            } else {
                if (ok0) { if (beta != 0.0) v0 += beta * cp[0]; cp[0] = v0; }
                if (ok1) { if (beta != 0.0) v1 += beta * cp[1]; cp[1] = v1; }
            }
        }
    }
}

inline int gemm_f64(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return 0;
    if (g.splitk > 1 && (g.batch != 1 || g.beta != 1.0)) return set_error(-3, "split-k GEMM needs batch == 1 and beta == 1");
    dim3 grid((unsigned)cdiv(g.M, GBM), (unsigned)cdiv(g.N, GBN), (unsigned)(g.splitk > 1 ? g.splitk : g.batch));
    if (grid.y > 65535) return set_error(-3, "too many column tiles");
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        auto setup = [](const void* f) {
            cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
            cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // room for GEMM_CTAS_PER_SM CTAs
        };
        setup((const void*)gemm_f64_kernel<0, 0>);
        setup((const void*)gemm_f64_kernel<0, 1>);
        setup((const void*)gemm_f64_kernel<1, 0>);
        setup((const void*)gemm_f64_kernel<1, 1>);
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
