// Evaluation bundle: the posterior-predictive samples and their 95 % interval coverage, which the
// reference's `performance_metrics` (trainers_regression.py:330-333, 181-224) obtains by pushing a 100x-repeated X through
// the whole q(f) computation (`sample_from_predictive_distribution`, sparse_MF_SP.py:886-992) and then calling numpy.quantile
// on the host.  Here the marginals (mu, v) of the test-NLL pass are reused: one warp per row draws S <= 128 samples
//   f_s = mu + sqrt(v) eps_s,   y_s = G(f_s) + sigma eta_s          (eps, eta ~ N(0,1): Philox-4x32-10 + Box-Muller)
// sorts them across the warp (bitonic, four values per lane) and forms numpy.quantile's linearly interpolated 2.5 % / 97.5 %
// quantiles, the coverage indicator of y_n and — optionally — exports the samples (tests compare the quantiles with
// numpy.quantile on exactly those samples).
#pragma once
#include "row_kernels.cuh"
#include "flow_mlp.cuh"       // philox4x32_10

namespace tgp {

struct RowCoverArgs {
    int R, n_rowp, n_mc, S;
    const double *mu, *v, *y, *log_var_noise, *theta, *rowp;   // rowp: (R, n_mc, n_rowp), n_mc in {1, S}
    unsigned long long seed;
    const unsigned long long* offset_dev;
    double q_lo_p, q_hi_p;                  // quantile levels (0.025, 0.975)
    double *q_lo, *q_hi, *covered;          // (R) each; covered = 1.0 / 0.0
    double* samples;                        // optional (R, S)
    double* count;                          // optional: += number of covered rows
    FlowDesc flow;
};

__device__ __forceinline__ void cmpswap(double& a, double& b, bool up) {
    const bool sw = (a > b) == up;
    const double t = a;
    a = sw ? b : a;
    b = sw ? t : b;
}

__global__ void __launch_bounds__(ROW_THREADS) k_row_coverage(const RowCoverArgs a) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    __shared__ double prep[FLOW_PREP_DOUBLES];
    flow_prepare(a.flow, a.theta, prep);
    const double sd_noise = sqrt(exp(a.log_var_noise[0]));
    const unsigned long long offset = a.offset_dev ? a.offset_dev[0] : 0ull;
    double local_count = 0.0;
    for (long n = (long)blockIdx.x * wpb + wid; n < a.R; n += (long)gridDim.x * wpb) {
        const double mu = a.mu[n], sd = sqrt(fmax(a.v[n], 0.0));
        double x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int s = lane + 32 * j;
            double val = INFINITY;                               // padding sorts to the end
            if (s < a.S) {
                uint32_t r[4];
                philox4x32_10((uint32_t)n, (uint32_t)((unsigned long long)n >> 32), (uint32_t)s, 0x5eedu, (uint32_t)a.seed ^ (uint32_t)offset,
                              (uint32_t)(a.seed >> 32) ^ (uint32_t)(offset >> 32), r);
                // two uniforms in (0, 1] from 2 x 32 bits each -> Box-Muller pair
                const double u1 = ((double)r[0] * 4294967296.0 + (double)r[1] + 1.0) * 5.421010862427522e-20;
                const double u2 = ((double)r[2] * 4294967296.0 + (double)r[3] + 1.0) * 5.421010862427522e-20;
                const double rad = sqrt(-2.0 * log(u1));
                double sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
                const double f = mu + sd * (rad * cs);
                const double* rowp = a.rowp ? a.rowp + ((long)n * a.n_mc + (a.n_mc > 1 ? s : 0)) * a.n_rowp : nullptr;
                double dG;
                val = flow_forward(a.flow, f, a.theta, rowp, &dG, nullptr, nullptr, prep) + sd_noise * (rad * sn);
                if (a.samples) a.samples[(long)n * a.S + s] = val;
            }
            x[j] = val;
        }
        // bitonic sort of the 128 values, element index i = 32 j + lane (ascending): strides < 32 exchange across lanes,
        // strides >= 32 inside the lane
        for (int k = 2; k <= 128; k <<= 1) {
            for (int st = k >> 1; st > 0; st >>= 1) {
                if (st == 64) {                                   // pairs (0,2), (1,3): indices 32 j + lane, k = 128 -> ascending
                    cmpswap(x[0], x[2], true);
                    cmpswap(x[1], x[3], true);
                } else if (st == 32) {                            // pairs (0,1), (2,3)
                    cmpswap(x[0], x[1], ((lane) & k) == 0);
                    cmpswap(x[2], x[3], ((64 + lane) & k) == 0);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = 32 * j + lane;
                        const double other = __shfl_xor_sync(0xffffffffu, x[j], st);
                        const bool up = (i & k) == 0;
                        const bool lower = (lane & st) == 0;          // this lane holds the smaller index of the pair
                        const bool take_min = lower == up;
                        x[j] = take_min ? fmin(x[j], other) : fmax(x[j], other);
                    }
                }
            }
        }
        // numpy.quantile (linear): h = q (S - 1); value = x[floor h] + (h - floor h) (x[floor h + 1] - x[floor h])
        auto pick = [&](int idx) -> double {
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((idx >> 5) == j) v = x[j];
            return __shfl_sync(0xffffffffu, v, idx & 31);
        };
        auto quant = [&](double q) -> double {
            const double h = q * (double)(a.S - 1);
            const int lo = (int)floor(h);
            const int hi = min(lo + 1, a.S - 1);
            const double xl = pick(lo), xh = pick(hi);
            return xl + (h - (double)lo) * (xh - xl);
        };
        const double ql = quant(a.q_lo_p), qh = quant(a.q_hi_p);
        if (lane == 0) {
            const double y = a.y[n];
            const double c = (y >= ql && y <= qh) ? 1.0 : 0.0;
            a.q_lo[n] = ql; a.q_hi[n] = qh; a.covered[n] = c;
            local_count += c;
        }
    }
    if (a.count) {
        local_count = warp_sum(local_count);
        if (lane == 0 && local_count != 0.0) atomicAdd(a.count, local_count);
    }
}

}  // namespace tgp
