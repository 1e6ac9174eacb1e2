// Tensor-core GEMM for the FP32 mode: 3xTF32 error-compensated product on tcgen05 with TMEM accumulators.
//
//   D[m,n] (+)= sum_k A[m,k] * B[n,k]        A: (Mrows x K), B: (Ncols x K), both K-major FP32 in global memory
//
// Every operand exists as two planes: `x` (raw FP32; the tensor core reads its top 19 bits = TF32 truncation) and
// `x_lo = x - tf32_trunc(x)` (exact).  Three MMAs per k-step accumulate  A*B + A_lo*B + A*B_lo  in FP32 in TMEM,
// which recovers ~FP32 accuracy (the dropped lo*lo term is 2^-22 relative).
//
// Persistent, warp-specialised CTA (192 threads), one CTA per SM:
//   warp 0    : TMA producer — cp.async.bulk.tensor (128B swizzle) of the four operand planes into a 2-stage ring
//   warp 1    : TMEM allocation + single-thread tcgen05.mma issue; tcgen05.commit releases stages / signals tiles
//   warps 2-9 : epilogue — tcgen05.ld 32 lanes x 32 columns at a time, FP32 register accumulation over k-chunks,
//               then FP32 store or FP64 atomic accumulation
// Tile 128 x 256 x 32 (BK = 32 floats = one 128-byte swizzle row); two 256-column TMEM accumulators (all 512 columns)
// so that the epilogue of tile t overlaps the MMAs of tile t+1.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_f64.cuh"      // GemmTimer (shared live-timing hook)

namespace tgp {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4;              // 16 KiB
constexpr int B_BYTES = BN * BK * 4;              // 32 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int EPI_WARPS = 8;                      // two warps per TMEM lane quarter, 128 columns each
constexpr int THREADS = 64 + 32 * EPI_WARPS;      // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int UMMA_K = 8;                         // tf32: 32 bytes per MMA k-step
// The tensor core accumulates into TMEM with truncation, so a 1024-deep FP32 accumulation drifts by ~1e-5 relative.
// Accumulation is therefore two-level: at most KCHUNK k-elements are summed in TMEM, the partial tiles are then added in
// round-to-nearest FP32 registers by the epilogue warps (which also lets the reduction be arbitrarily deep).
#ifndef TGP_KCHUNK
#define TGP_KCHUNK 128
#endif
constexpr int KCHUNK = TGP_KCHUNK;
constexpr int KB_PER_CHUNK = KCHUNK / BK;

struct Params {
    int Mrows, Ncols, K;
    // k-range clipping from triangular structure (all in elements, multiples of BK are derived inside):
    //   tri_mode 0: dense
    //   tri_mode 1: B rows n < tri_rows hold a lower-triangular block (nonzero k <= n): k_end = min(K, n0 + BN)
    //   tri_mode 2: k in [0, tri_rows) is "upper" w.r.t. n (nonzero k >= n): k_begin = n0       (k >= tri_rows dense)
    int tri_mode, tri_rows;
    // output: mode 0: Cf[m*ldc + n] = acc (FP32);  mode 1: atomicAdd(Cd[m*ldc + n], acc) (FP64)
    // lower_rows > 0 (mode 1): for m < lower_rows only n <= m is written and tiles strictly above the diagonal skipped
    int out_mode, lower_rows;
    float* Cf; double* Cd; long ldc;
    int splitk;                                   // > 1: grid-stride tiles are (tile, split) pairs; out_mode must be 1
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, 128-byte swizzle: rows are 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1
__device__ __forceinline__ uint64_t make_desc(const void* smem_tile) {
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, FP32 accumulate, both operands K-major, M = 128, N = BN
__device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void k_range(const Params& p, int m0, int n0, int& kb, int& ke) {
    kb = 0; ke = p.K;
    if (p.tri_mode == 1 && n0 < p.tri_rows) ke = min(p.K, n0 + BN);
    if (p.tri_mode == 2 && n0 < p.tri_rows) kb = (n0 / BK) * BK;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapBlo, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;        // [2] accumulator ready
    uint64_t* tempty = tfull + 2;            // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.Mrows + BM - 1) / BM, tiles_n = (p.Ncols + BN - 1) / BN;
    const int splitk = p.splitk > 1 ? p.splitk : 1;
    const long n_work = (long)tiles_m * tiles_n * splitk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // decode work item -> (m-tile, n-tile, split); n fastest so that consecutive CTAs share the A row block
    auto decode = [&](long w, int& m0, int& n0, int& kb, int& ke) -> bool {
        const int sp = (int)(w % splitk);
        const long t = w / splitk;
        n0 = (int)(t % tiles_n) * BN;
        m0 = (int)(t / tiles_n) * BM;
        if (p.lower_rows > 0 && m0 < p.lower_rows && n0 > m0 + BM - 1) return false;
        k_range(p, m0, n0, kb, ke);
        int nk = ke > kb ? (ke - kb + BK - 1) / BK : 0;
        if (splitk > 1) {
            const int per = (nk + splitk - 1) / splitk;
            const int t0 = min(nk, sp * per), t1 = min(nk, t0 + per);
            kb += t0 * BK;
            ke = min(ke, kb + (t1 - t0) * BK);
            nk = t1 - t0;
        }
        return nk > 0;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
                int m0, n0, kb, ke;
                if (!decode(w, m0, n0, kb, ke)) continue;
                for (int k = kb; k < ke; k += BK) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(st, &mapA, &full[stage], k, m0);
                    tma_load_2d(st + A_BYTES, &mapAlo, &full[stage], k, m0);
                    tma_load_2d(st + 2 * A_BYTES, &mapB, &full[stage], k, n0);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &mapBlo, &full[stage], k, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t bphase = 0;
            for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
                int m0, n0, kb, ke;
                if (!decode(w, m0, n0, kb, ke)) continue;
                int kbi = 0;                                  // k-block index inside the current chunk
                uint32_t tmem_d = 0, accum = 0;
                for (int k = kb; k < ke; k += BK) {
                    if (kbi == 0) {                            // new accumulation chunk -> next TMEM buffer
                        mbar_wait(&tempty[buf], bphase ^ 1);
                        tc_fence_after();
                        tmem_d = tmem_base + (uint32_t)buf * BN;
                        accum = 0;
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    const uint64_t dA = make_desc(st), dAl = make_desc(st + A_BYTES);
                    const uint64_t dB = make_desc(st + 2 * A_BYTES), dBl = make_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);      // +32 B per k-step inside the swizzle row
                        umma_tf32(tmem_d, dA + adv, dB + adv, idesc, accum);
                        accum = 1;
                        umma_tf32(tmem_d, dAl + adv, dB + adv, idesc, 1);
                        umma_tf32(tmem_d, dA + adv, dBl + adv, idesc, 1);
                    }
                    umma_commit(&empty[stage]);               // frees the smem stage when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    if (++kbi == KB_PER_CHUNK || k + BK >= ke) {
                        umma_commit(&tfull[buf]);             // chunk accumulator complete
                        if (++buf == 2) { buf = 0; bphase ^= 1; }
                        kbi = 0;
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps 2..9: TMEM lanes 32*(warp%4) .. +31, columns 128*half .. +127 =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int buf = 0; uint32_t bphase = 0;
        for (long w = blockIdx.x; w < n_work; w += gridDim.x) {
            int m0, n0, kb, ke;
            if (!decode(w, m0, n0, kb, ke)) continue;
            const int nchunks = ((ke - kb + BK - 1) / BK + KB_PER_CHUNK - 1) / KB_PER_CHUNK;
            float acc[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) acc[i] = 0.f;
            for (int ch = 0; ch < nchunks; ++ch) {
                mbar_wait(&tfull[buf], bphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (uint32_t)buf * BN + (uint32_t)(half * 128) + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int c = 0; c < 128; c += 32) {
                    uint32_t r[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr + (uint32_t)c));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c + i] += __uint_as_float(r[i]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
            const int row = m0 + q * 32 + lane;
            if (row < p.Mrows) {
                const int nbase = n0 + half * 128;
                if (p.out_mode == 0) {
                    float* dst = p.Cf + (long)row * p.ldc + nbase;
                    if (nbase + 128 <= p.Ncols && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                        for (int i = 0; i < 128; i += 4)
                            *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 128; ++i) if (nbase + i < p.Ncols) dst[i] = acc[i];
                    }
                } else {
                    double* dst = p.Cd + (long)row * p.ldc + nbase;
                    const bool lower = p.lower_rows > 0 && row < p.lower_rows;
#pragma unroll
                    for (int i = 0; i < 128; ++i)
                        if (nbase + i < p.Ncols && (!lower || nbase + i <= row)) atomicAdd(dst + i, (double)acc[i]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2D K-major FP32 matrix (rows x cols, leading dimension ld floats, ld % 4 == 0), box = 32 floats x box_rows, 128B swizzle.
// Encoded descriptors are cached (the step re-uses the same workspaces every iteration, so after the first step every
// launch finds its four maps here instead of calling into the driver).
struct MapKey { const float* base; long rows, cols, ld; int box_rows; };
struct MapCache {
    static constexpr int N = 64;
    MapKey key[N]; CUtensorMap map[N]; int used = 0, next = 0;
    const CUtensorMap* find(const MapKey& k) const {
        for (int i = 0; i < used; ++i)
            if (key[i].base == k.base && key[i].rows == k.rows && key[i].cols == k.cols && key[i].ld == k.ld && key[i].box_rows == k.box_rows)
                return &map[i];
        return nullptr;
    }
    void put(const MapKey& k, const CUtensorMap& m) {
        const int i = used < N ? used++ : (next = (next + 1) % N);
        key[i] = k; map[i] = m;
    }
};
inline MapCache& map_cache() { static MapCache c; return c; }

inline int make_map(CUtensorMap* map, const float* base, long rows, long cols, long ld, int box_rows) {
    const MapKey mk{base, rows, cols, ld, box_rows};
    if (const CUtensorMap* hit = map_cache().find(mk)) { *map = *hit; return 0; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(-101, "cuTensorMapEncodeTiled not available from the driver");
    if ((ld & 3) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(-2, "TMA operand must be 16-byte aligned with ld % 4 == 0");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(-102, "cuTensorMapEncodeTiled failed");
    map_cache().put(mk, *map);
    return 0;
}

struct Operand { const float* hi; const float* lo; long rows, cols, ld; };

inline int gemm_tf32x3(const Operand& A, const Operand& B, Params p, cudaStream_t st) {
    if (p.Mrows <= 0 || p.Ncols <= 0 || p.K <= 0) return 0;
    CUtensorMap mA, mAl, mB, mBl;
    TGP_TRY(make_map(&mA, A.hi, A.rows, A.cols, A.ld, BM));
    TGP_TRY(make_map(&mAl, A.lo, A.rows, A.cols, A.ld, BM));
    TGP_TRY(make_map(&mB, B.hi, B.rows, B.cols, B.ld, BN));
    TGP_TRY(make_map(&mBl, B.lo, B.rows, B.cols, B.ld, BN));
    static PerDeviceOnce attr_once;
    if (attr_once.first()) cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    const long tiles = (long)cdiv(p.Mrows, BM) * cdiv(p.Ncols, BN) * (p.splitk > 1 ? p.splitk : 1);
    int sms = 148;
    const int grid = (int)(tiles < sms ? tiles : sms);
    const bool timed = g_gemm_timer.enabled;
    if (timed) g_gemm_timer.begin(2, st);
    gemm_tf32x3_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(mA, mAl, mB, mBl, p);
    if (timed) g_gemm_timer.end(st);
    return check_launch("gemm_tf32x3_kernel");
}

}  // namespace tc
}  // namespace tgp
