// Gauss-Seidel post-process of the reference (its `if 0:` branch, optimize_looper.py:55-74):
// misc/optimize_loop.py::optimize_loop (4-neighbour smoothing) and
// misc/opt_loop.py::make_weight / optimize_loop_bilateral_horizon / _vertical.
//
// The reference visits the cells one by one and updates the image IN PLACE, so a cell sees the new
// values of the cells visited before it.  That order is kept -- only the cells that cannot see each
// other are computed side by side:
//   * 4-neighbour forward sweep: the cells of an anti-diagonal i + j = t;
//   * 4-neighbour "reverse" sweep: the reference flips the row index inside the inner loop, so even
//     visits walk their column bottom-up and odd visits top-down; all columns advance in lockstep and
//     the one step where they meet on the same row is walked sequentially (oracle/dm_oracle.py says why);
//   * bilateral sweep over a (2e+1)^2 window: the cells with the same j + (e+1) i.
// One CTA walks the whole image (a block barrier per step); float64, no contraction, numpy's pairwise
// summation order for the window sums: bit-identical to the reference given the same inputs.
#include "dm_common.cuh"

namespace {

constexpr int GS_THREADS = 1024;

// np.where(arr > hi, hi, arr); np.where(arr < lo, lo, arr): NaN stays
__device__ __forceinline__ double clamp_like_where(double v, double lo, double hi) {
    if (v > hi) v = hi;
    if (v < lo) v = lo;
    return v;
}

// sequential sum of n doubles in index order by ONE thread (error += ...), staged through shared memory
__device__ double sequential_sum(const double* __restrict__ v, long long n, double* stage /* [GS_THREADS] shared */) {
    double acc = 0.0;
    for (long long base = 0; base < n; base += GS_THREADS) {
        const long long idx = base + threadIdx.x;
        stage[threadIdx.x] = idx < n ? v[idx] : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int m = (int)((n - base) < GS_THREADS ? (n - base) : GS_THREADS);
            for (int k = 0; k < m; ++k) acc = __dadd_rn(acc, stage[k]);
        }
        __syncthreads();
    }
    return acc;         // valid in thread 0
}

// misc/optimize_loop.py:15-37
__global__ void __launch_bounds__(GS_THREADS)
dm_optimize_loop_kernel(volatile double* d, const double* __restrict__ co, int S0, int S1, int e, double alpha,
                        double* __restrict__ diff, double* __restrict__ error_out) {
    __shared__ double stage[GS_THREADS];
    for (long long idx = threadIdx.x; idx < (long long)S0 * S1; idx += GS_THREADS) d[idx] = clamp_like_where(d[idx], 0.0, 10.0);
    __syncthreads();
    const int lo0 = e, hi0 = S0 - e - 1, lo1 = e, hi1 = S1 - e - 1;        // half-open
    if (hi0 <= lo0 || hi1 <= lo1) { if (threadIdx.x == 0) *error_out = 0.0; return; }
    auto update = [&](int i, int j) -> double {
        const size_t c = (size_t)i * S1 + j;
        const double sum_d = __dadd_rn(__dadd_rn(__dadd_rn(d[c - 1], d[c + 1]), d[c - S1]), d[c + S1]);
        const double a = co[c];
        return __ddiv_rn(__dadd_rn(__dmul_rn(-a, d[c]), __dmul_rn(alpha, sum_d)), __dadd_rn(-a, __dmul_rn(4.0, alpha)));
    };
    for (int t = lo0 + lo1; t <= hi0 + hi1 - 2; ++t) {
        const int ia = max(lo0, t - (hi1 - 1)), ib = min(hi0 - 1, t - lo1);
        for (int i = ia + threadIdx.x; i <= ib; i += GS_THREADS) {
            const double v = update(i, t - i);
            d[(size_t)i * S1 + (t - i)] = v;
        }
        __syncthreads();
    }
    const int R = S0 - lo0 - 1, C = S1 - lo1 - 1, n0 = hi0 - lo0, n1 = hi1 - lo1;
    for (int o = 0; o < n0; ++o) {
        if (2 * o == R - lo0) {
            if (threadIdx.x == 0)
                for (int k = 0; k < n1; ++k) {
                    const int r = (k & 1) ? lo0 + o : R - o, c = C - k;
                    const double v = update(r, c);
                    diff[(size_t)o * n1 + k] = fabs(__dsub_rn(d[(size_t)r * S1 + c], v));
                    d[(size_t)r * S1 + c] = v;
                }
        } else {
            for (int k = threadIdx.x; k < n1; k += GS_THREADS) {
                const int r = (k & 1) ? lo0 + o : R - o, c = C - k;
                const double v = update(r, c);
                diff[(size_t)o * n1 + k] = fabs(__dsub_rn(d[(size_t)r * S1 + c], v));
                d[(size_t)r * S1 + c] = v;
            }
        }
        __syncthreads();
    }
    const double err = sequential_sum(diff, (long long)n0 * n1, stage);
    if (threadIdx.x == 0) *error_out = err;
}

// misc/opt_loop.py:66-85
__global__ void dm_make_weight_kernel(const double* __restrict__ guide, int S0, int S1, int e, double sigma0, double sigma1,
                                      double* __restrict__ gw, double* __restrict__ cw) {
    const int w = 2 * e + 1, ww = w * w;
    const long long total = (long long)(S0 - e) * (S1 - e) * ww;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ww) {
        const int a = (int)idx / w - e, b = (int)idx % w - e;
        gw[idx] = exp(__ddiv_rn(-(double)(a * a + b * b), __dmul_rn(2.0, __dmul_rn(sigma1, sigma1))));
    }
    if (idx >= total) return;
    const int k = (int)(idx % ww);
    const long long cell = idx / ww;
    const int cj = (int)(cell % (S1 - e)), ci = (int)(cell / (S1 - e));      // ci = i - e
    const int i = ci + e, j = cj + e;
    double v = 0.0;                                                          // rows / columns the loops do not reach stay zero
    if (i < S0 - e - 1 && j < S1 - e - 1) {
        const double c = __dsub_rn(guide[(size_t)i * S1 + j], guide[(size_t)(i - e + k / w) * S1 + (j - e + k % w)]);
        v = exp(__ddiv_rn(__dmul_rn(__dmul_rn(-1.0, c), c), __dmul_rn(2.0, __dmul_rn(sigma0, sigma0))));
    }
    cw[idx] = v;
}

// numpy's pairwise summation of n contiguous doubles produced by f(k) (numpy/core/src/umath/loops_utils.h:
// pairwise_sum, block size 128, eight accumulators)
template <typename F>
__device__ double pairwise_sum(F f, int first, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int k = 0; k < n; ++k) res = __dadd_rn(res, f(first + k));
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = f(first + k);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], f(first + i + k));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, f(first + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(pairwise_sum(f, first, n2), pairwise_sum(f, first + n2, n - n2));
}

// misc/opt_loop.py:16-63
__global__ void __launch_bounds__(GS_THREADS)
dm_optimize_loop_bilateral_kernel(volatile double* d, const double* __restrict__ cw, const double* __restrict__ gw,
                                  const double* __restrict__ co, int S0, int S1, int e, int vertical,
                                  double* __restrict__ diff, double* __restrict__ error_out) {
    __shared__ double stage[GS_THREADS];
    const int lo0 = e, hi0 = S0 - e - 1, lo1 = e, hi1 = S1 - e - 1;
    if (hi0 <= lo0 || hi1 <= lo1) { if (threadIdx.x == 0) *error_out = 0.0; return; }
    const int w = 2 * e + 1, ww = w * w, n0 = hi0 - lo0, n1 = hi1 - lo1;
    // the reference reads CONSTANT entries of the coefficient array here (opt_loop.py:30-31,52-53)
    const double c0 = co[(size_t)e * S1 + e];
    const double cp = vertical ? co[(size_t)(e + 1) * S1 + e] : co[(size_t)e * S1 + e + 1];
    const double cm = vertical ? co[(size_t)(e - 1) * S1 + e] : co[(size_t)e * S1 + e - 1];
    const double a = -__dsub_rn(c0, __ddiv_rn(__dadd_rn(cp, cm), 2.0));
    const double shift = __ddiv_rn(__ddiv_rn(__dsub_rn(cp, cm), 2.0), __dadd_rn(__dadd_rn(__dmul_rn(-2.0, c0), cp), cm));
    for (int t = lo1 + (e + 1) * lo0; t <= (hi1 - 1) + (e + 1) * (hi0 - 1); ++t) {
        for (int i = lo0 + threadIdx.x; i < hi0; i += GS_THREADS) {
            const int j = t - (e + 1) * i;
            if (j < lo1 || j >= hi1) continue;
            const double* cwc = cw + ((size_t)(i - e) * (S1 - e) + (j - e)) * ww;
            const size_t org = (size_t)(i - e) * S1 + (j - e);
            auto weight = [&](int k) -> double { return __dmul_rn(gw[k], cwc[k]); };
            auto weighted = [&](int k) -> double { return __dmul_rn(weight(k), d[org + (size_t)(k / w) * S1 + (k % w)]); };
            const double num = pairwise_sum(weighted, 0, ww), den = pairwise_sum(weight, 0, ww);
            const double cur = d[(size_t)i * S1 + j];
            const double b = __dsub_rn(cur, shift);
            const double v = __ddiv_rn(__dadd_rn(__dmul_rn(-a, b), num), __dadd_rn(-a, den));
            diff[(size_t)(i - lo0) * n1 + (j - lo1)] = fabs(__dsub_rn(cur, v));
            d[(size_t)i * S1 + j] = v;
        }
        __syncthreads();
    }
    const double err = sequential_sum(diff, (long long)n0 * n1, stage);
    if (threadIdx.x == 0) *error_out = err;
}

}  // namespace

extern "C" int dm_optimize_loop(double* img_dev, const double* coef_dev, int s0, int s1, int exclusion, double alpha,
                                double* diff_dev, double* error_dev, void* stream) {
    DM_REQUIRE(img_dev && coef_dev && diff_dev && error_dev && s0 > 0 && s1 > 0 && exclusion >= 1, DM_ERR_INVALID, "dm_optimize_loop: bad arguments");
    dm_optimize_loop_kernel<<<1, GS_THREADS, 0, (cudaStream_t)stream>>>(img_dev, coef_dev, s0, s1, exclusion, alpha, diff_dev, error_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_make_weight(const double* guide_dev, int s0, int s1, int exclusion, double sigma0, double sigma1,
                              double* gw_dev, double* cw_dev, void* stream) {
    DM_REQUIRE(guide_dev && gw_dev && cw_dev && exclusion >= 1 && s0 > exclusion && s1 > exclusion, DM_ERR_INVALID, "dm_make_weight: bad arguments");
    const int w = 2 * exclusion + 1;
    const long long total = (long long)(s0 - exclusion) * (s1 - exclusion) * w * w;
    dm_make_weight_kernel<<<dm_div_up(total, 256), 256, 0, (cudaStream_t)stream>>>(guide_dev, s0, s1, exclusion, sigma0, sigma1, gw_dev, cw_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_optimize_loop_bilateral(double* img_dev, const double* cw_dev, const double* gw_dev, const double* coef_dev,
                                          int s0, int s1, int exclusion, int vertical, double* diff_dev, double* error_dev, void* stream) {
    DM_REQUIRE(img_dev && cw_dev && gw_dev && coef_dev && diff_dev && error_dev && exclusion >= 1, DM_ERR_INVALID, "dm_optimize_loop_bilateral: bad arguments");
    DM_REQUIRE(s0 > exclusion + 1 && s1 > exclusion + 1, DM_ERR_INVALID, "dm_optimize_loop_bilateral: image %dx%d too small for exclusion %d", s0, s1, exclusion);
    dm_optimize_loop_bilateral_kernel<<<1, GS_THREADS, 0, (cudaStream_t)stream>>>(img_dev, cw_dev, gw_dev, coef_dev, s0, s1, exclusion, vertical != 0, diff_dev, error_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}
