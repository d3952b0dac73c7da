// Top-down backtracking and the level-0 parabola refinement.
// Replaces Matching._calc_near_match (misc/Matching.py:58-78), _initial_move_map
// (:80-96), _B (:98-139), the displacement filter _filter (:224-255),
// _sub_pix_compute/_sub_pix_cal (:165-209) and the assembly of the (3,T0,T1) float64 result
// (:211-222).
//
// Index results are bit-exact with the reference given the same level data: only
// comparisons and one addition are involved, and the kernels are instantiated for both
// float (the pipeline's dtype) and double (the reference's dtype).
#include "dm_common.cuh"

namespace {

template <typename T> struct Near { static __device__ __forceinline__ T zero_thr(); };
template <> struct Near<float>  { static __device__ __forceinline__ float  zero_thr() { return DM_NEAR_ZERO_F; } };
template <> struct Near<double> { static __device__ __forceinline__ double zero_thr() { return DM_NEAR_ZERO_D; } };

// misc/Matching.py:58-78.  map = one (C,D) slice; (d0,d1) = p_dot.
// np.argmax: first maximum in row-major order, the first NaN wins; zero padding outside.
template <typename T>
__device__ __forceinline__ void near_match(const T* __restrict__ map, int C, int D, int d0, int d1,
                                           int& r0, int& r1, T& score) {
    T best = T(0);
    int bi = 0;
    bool best_nan = false;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int y = d0 + k / 3 - 1, x = d1 + k % 3 - 1;
        T v = T(0);
        if (y >= 0 && y < C && x >= 0 && x < D) v = map[(size_t)y * D + x];
        if (k == 0) { best = v; best_nan = (v != v); }
        else if (!best_nan && (v > best || v != v)) { best = v; bi = k; best_nan = (v != v); }
    }
    T centre = T(0);
    if (d0 >= 0 && d0 < C && d1 >= 0 && d1 < D) centre = map[(size_t)d0 * D + d1];
    if (best < Near<T>::zero_thr()) { bi = 4; best = centre; }      // false for NaN, like numpy
    r0 = d0 + bi / 3 - 1;
    r1 = d1 + bi % 3 - 1;
    score = best + centre;
}

template <typename T>
__global__ void __launch_bounds__(256)
dm_backtrack_kernel(const T* __restrict__ level, long long total, int A, int B, int C, int D,
                    const int32_t* __restrict__ parent, int32_t* __restrict__ match, T* __restrict__ score) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % B);
    long long t = idx / B;
    const int i = (int)(t % A);
    const long long n = t / A;
    int d0, d1;
    if (parent) {
        const int hA = A >> 1, hB = B >> 1;
        const size_t pbase = (size_t)n * 2 * hA * hB;
        const size_t pidx = (size_t)(i >> 1) * hB + (j >> 1);
        d0 = 2 * parent[pbase + pidx] + (i & 1);
        d1 = 2 * parent[pbase + (size_t)hA * hB + pidx] + (j & 1);
    } else {
        d0 = i; d1 = j;
    }
    const T* map = level + (((size_t)n * A + i) * B + j) * (size_t)C * D;
    int r0, r1;
    T s;
    near_match<T>(map, C, D, d0, d1, r0, r1, s);
    const size_t mbase = (size_t)n * 2 * A * B;
    const size_t midx = (size_t)i * B + j;
    match[mbase + midx] = r0;
    match[mbase + (size_t)A * B + midx] = r1;
    score[(size_t)n * A * B + midx] = s;
}

// misc/Matching.py:165-175
template <typename T>
__device__ __forceinline__ T sub_pix_fit(T r0, T r1, T rm) {
    if (r0 > r1 && r0 > rm) return -(r1 - rm) / (T(2) * (r1 + rm - T(2) * r0));
    return T(0);
}

// WIDE: float32 storage of the library's own pyramid, differences formed in float64 -- the
// reference holds those values in float64 arrays (dm_parabola_shift); otherwise the arithmetic is
// the storage type's, as numpy does for a float32 / float64 co_map_list handed in by the caller.
template <typename T, bool WIDE>
__global__ void __launch_bounds__(256)
dm_match_map_kernel(const T* __restrict__ l0, long long total, int T0, int T1,
                    const int32_t* __restrict__ match, const T* __restrict__ score, int sub_pix,
                    double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int P = T0 * T1;
    const long long n = idx / P;
    const int p = (int)(idx - n * P);
    const int c0 = match[(size_t)n * 2 * P + p];
    const int c1 = match[(size_t)n * 2 * P + P + p];
    double m0 = (double)c0, m1 = (double)c1;
    // numpy index rules (misc/Matching.py:186-208): -size <= index < size is accepted and a
    // negative one wraps; anything else raises IndexError, which the reference swallows ->
    // that axis stays unrefined.  Matches only leave the map when the displacement filter
    // ran on level 0.
    if (sub_pix && dm_np_index_ok(c0, T0) && dm_np_index_ok(c1, T1)) {
        const T* map = l0 + ((size_t)n * P + p) * (size_t)P;    // (C,D) = (T0,T1)
        const int w0 = dm_np_wrap(c0, T0), w1 = dm_np_wrap(c1, T1);
        const T r0 = map[(size_t)w0 * T1 + w1];
        if (dm_np_index_ok(c0 + 1, T0) && dm_np_index_ok(c0 - 1, T0)) {
            const T r1 = map[(size_t)dm_np_wrap(c0 + 1, T0) * T1 + w1];
            const T rm = map[(size_t)dm_np_wrap(c0 - 1, T0) * T1 + w1];
            m0 += WIDE ? ((r0 > r1 && r0 > rm) ? dm_parabola_shift((float)r0, (float)r1, (float)rm) : 0.0) : (double)sub_pix_fit<T>(r0, r1, rm);
        }
        if (dm_np_index_ok(c1 + 1, T1) && dm_np_index_ok(c1 - 1, T1)) {
            const T r1 = map[(size_t)w0 * T1 + dm_np_wrap(c1 + 1, T1)];
            const T rm = map[(size_t)w0 * T1 + dm_np_wrap(c1 - 1, T1)];
            m1 += WIDE ? ((r0 > r1 && r0 > rm) ? dm_parabola_shift((float)r0, (float)r1, (float)rm) : 0.0) : (double)sub_pix_fit<T>(r0, r1, rm);
        }
    }
    double* o = out + (size_t)n * 3 * P;
    o[p] = m0;
    o[P + p] = m1;
    o[2 * P + p] = (double)score[(size_t)n * P + p];
}

// misc/Matching.py:224-255 (_filter): every interior cell of the (h,w) displacement field is
// replaced by round(mean | median of its (2e+1)^2 neighbourhood) + its own coordinate.  The
// neighbourhood is read from the unfiltered field (the reference snapshots it in d_map),
// Python's round() is half-to-even (rint), the median of an odd count is its middle element.
// Border cells are copied.  in / out: int32 [n][2][H][W].
constexpr int DM_FILTER_MAX_E = 4;      // windows up to 9 x 9 are ranked in registers, larger ones by bisection
__global__ void __launch_bounds__(256)
dm_match_filter_kernel(const int32_t* __restrict__ in, long long total, int H, int W, int e, int mode, int32_t* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j = (int)(idx % W);
    const long long t = idx / W;
    const int i = (int)(t % H);
    const long long n = t / H;
    const size_t HW = (size_t)H * W, base = (size_t)n * 2 * HW, cell = (size_t)i * W + j;
    const int32_t* rowp = in + base;        // matched row of every cell
    const int32_t* colp = in + base + HW;   // matched column
    int o0 = rowp[cell], o1 = colp[cell];
    if (i >= e && i < H - e && j >= e && j < W - e) {
        const int k = 2 * e + 1, cnt = k * k;
        if (mode == DM_FILTER_AVERAGE) {
            long long s0 = 0, s1 = 0;
            for (int y = i - e; y <= i + e; ++y)
                for (int x = j - e; x <= j + e; ++x) { s0 += rowp[(size_t)y * W + x] - y; s1 += colp[(size_t)y * W + x] - x; }
            o0 = (int)rint(__ddiv_rn((double)s0, (double)cnt)) + i;     // np.mean = sum / count in float64, then round()
            o1 = (int)rint(__ddiv_rn((double)s1, (double)cnt)) + j;
        } else if (e > DM_FILTER_MAX_E) {
            // large windows: the middle element is the smallest value v with count(x <= v) > cnt / 2;
            // bisection over the displacement range, the window is re-read from memory each time
            const int half = cnt / 2;
            int med[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int32_t* p = c == 0 ? rowp : colp;
                int lo = -(1 << 20), hi = 1 << 20;
                while (lo < hi) {
                    const int mid = lo + ((hi - lo) >> 1);
                    int le = 0;
                    for (int y = i - e; y <= i + e; ++y)
                        for (int x = j - e; x <= j + e; ++x) le += (p[(size_t)y * W + x] - (c == 0 ? y : x)) <= mid;
                    if (le > half) hi = mid; else lo = mid + 1;
                }
                med[c] = lo;
            }
            o0 = med[0] + i; o1 = med[1] + j;
        } else {
            int v0[(2 * DM_FILTER_MAX_E + 1) * (2 * DM_FILTER_MAX_E + 1)], v1[(2 * DM_FILTER_MAX_E + 1) * (2 * DM_FILTER_MAX_E + 1)];
            int m = 0;
            for (int y = i - e; y <= i + e; ++y)
                for (int x = j - e; x <= j + e; ++x) { v0[m] = rowp[(size_t)y * W + x] - y; v1[m] = colp[(size_t)y * W + x] - x; ++m; }
            // middle element by rank counting (cnt is odd): the value with <= cnt/2 smaller and > cnt/2 smaller-or-equal
            const int half = cnt / 2;
            int med0 = v0[0], med1 = v1[0];
            for (int a = 0; a < cnt; ++a) {
                int lt0 = 0, le0 = 0, lt1 = 0, le1 = 0;
                for (int b = 0; b < cnt; ++b) {
                    lt0 += v0[b] < v0[a]; le0 += v0[b] <= v0[a];
                    lt1 += v1[b] < v1[a]; le1 += v1[b] <= v1[a];
                }
                if (lt0 <= half && le0 > half) med0 = v0[a];
                if (lt1 <= half && le1 > half) med1 = v1[a];
            }
            o0 = med0 + i; o1 = med1 + j;
        }
    }
    out[base + cell] = o0;
    out[base + HW + cell] = o1;
}

// misc/Calc_difference.py:25-49
__global__ void dm_cal_map_kernel(const double* __restrict__ map, int T0, int T1, int mode, double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = T0 * T1;
    if (idx >= P) return;
    const int i = idx / T1, j = idx - i * T1;
    const double d0 = __dsub_rn((double)i, map[idx]);
    const double d1 = __dsub_rn((double)j, map[P + idx]);
    double v;
    if (mode == DM_MODE_ELEVATION) v = d1;
    else if (mode == DM_MODE_ELEVATION2) v = d0;
    else v = __dsqrt_rn(__fma_rn(d1, d1, __dmul_rn(d0, d0)));   // np.linalg.norm = sqrt(ddot) with the fused second product
    out[idx] = v;
}

// misc/sub_pix_cal.py:22-53 (+ misc/optimize_loop.py:40-44).  All float64, no contraction,
// so the result is bit-identical to numpy.
__device__ __forceinline__ double clamp3(double v) {
    if (v > 3.0) v = 3.0;          // np.where(arr > 3, 3, arr): NaN stays
    if (v < -3.0) v = -3.0;
    return v;
}
__global__ void dm_sub_pix_cal_kernel(const double* __restrict__ arr, const double* __restrict__ co,
                                      int S0, int S1, int direction, double ratio, double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S0 * S1) return;
    const int i = idx / S1, j = idx - i * S1;
    double d = clamp3(arr[idx]);
    double dis = d;
    if (i >= 1 && i < S0 - 1 && j >= 1 && j < S1 - 1) {
        const int step = direction == 0 ? S1 : 1;
        const double r0 = __dmul_rn(co[idx], ratio);
        const double r1 = __dmul_rn(co[idx + step], ratio);
        const double rm = __dmul_rn(co[idx - step], ratio);
        const double den = __dmul_rn(2.0, __dsub_rn(__dadd_rn(r1, rm), __dmul_rn(2.0, r0)));
        dis = __dsub_rn(d, __ddiv_rn(__dsub_rn(r1, rm), den));
        if (fabs(__dsub_rn(d, dis)) > 1.0) dis = d;
    }
    out[idx] = clamp3(dis);
}

// ---------------------------------------------------------------------------------------
// Upper pyramid tail + whole top-down pass of ONE tile in one CTA (fused scene solver, no
// displacement filter).  Level `ks` (at most 16 KiB per tile) has been built by dm_aggregate;
// the CTA builds the levels above it in shared memory (Correlation_map._aggregation,
// misc/Correlation_map.py:100-128) and then walks back down to level 1
// (Matching._initial_move_map / _B, misc/Matching.py:80-139) with the matches of the level above
// in shared memory.  Same arithmetic and tie-breaks as dm_aggregate_*_kernel and
// dm_backtrack_kernel; replaces ~9 small dependent launches.  (Building level ks here as well --
// 4096 outputs of 36 gathered loads each per CTA -- took 0.1 ms and lost to the generic kernel.)
struct TailArgs {
    float* level[16];           // global level buffers [n][A][B][C][D] (levels >= ks are also written)
    int t0, t1, L, ks;
    int32_t* match1;            // [n][2][t0/2][t1/2]: matches of level 1
};

__device__ __forceinline__ void tail_aggregate(const float* __restrict__ in, int A, int B, int C, int D,
                                               float* __restrict__ out_s, float* __restrict__ out_g) {
    const int hB = B >> 1, oc = C >> 1, od = D >> 1;
    const int total = (A >> 1) * hB * oc * od;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int x = idx % od;
        int t = idx / od;
        const int y = t % oc; t /= oc;
        const int J = t % hB, I = t / hB;
        float sum = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const float* src = in + ((size_t)(2 * I + (ch >> 1)) * B + (2 * J + (ch & 1))) * (size_t)(C * D);
            float m = -CUDART_INF_F;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int r = 2 * y + dy;                 // r < C always (C even)
                if (r < 0) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int c = 2 * x + dx;
                    if (c < 0) continue;
                    m = dm_max_nan(m, src[(size_t)r * D + c]);
                }
            }
            sum = (ch == 0) ? m : __fadd_rn(sum, m);      // ((ul + ur) + ll) + lr
        }
        const float o = dm_rectify(__fmul_rn(sum, 0.25f));
        out_s[idx] = o;
        out_g[idx] = o;
    }
}

__global__ void __launch_bounds__(256)
dm_upper_tail_kernel(const TailArgs a) {
    extern __shared__ __align__(16) float tail_smem[];
    dm_pdl_wait();                  // level ks comes from the aggregation kernel in front
    dm_pdl_launch_dependents();
    const int n = blockIdx.x, L = a.L, ks = a.ks;
    // shared levels ks+1 .. L-1, then two match buffers of [2][cells of level 2]
    float* slevel[16];
    size_t off = 0;
    for (int k = ks + 1; k < L; ++k) {
        slevel[k] = tail_smem + off;
        const size_t cells = (size_t)(a.t0 >> k) * (a.t1 >> k);
        off += cells * cells;
    }
    const int cells2 = (a.t0 >> 2) * (a.t1 >> 2);
    int32_t* mprev = reinterpret_cast<int32_t*>(tail_smem + off);
    int32_t* mcur = mprev + 2 * cells2;

    for (int k = ks; k + 1 < L; ++k) {                    // level k -> k + 1
        const int A = a.t0 >> k, B = a.t1 >> k;
        const size_t cells = (size_t)A * B, ncells = (size_t)(A >> 1) * (B >> 1);
        const float* in = (k > ks) ? slevel[k] : a.level[k] + (size_t)n * cells * cells;
        tail_aggregate(in, A, B, A, B, slevel[k + 1], a.level[k + 1] + (size_t)n * ncells * ncells);
        __syncthreads();
    }
    for (int k = L - 1; k >= 1; --k) {
        const int A = a.t0 >> k, B = a.t1 >> k, cells = A * B;
        const float* lv = (k > ks) ? slevel[k] : a.level[k] + (size_t)n * cells * cells;
        for (int idx = threadIdx.x; idx < cells; idx += blockDim.x) {
            const int i = idx / B, j = idx - i * B;
            int d0 = i, d1 = j;
            if (k < L - 1) {                              // misc/Matching.py:116-124
                const int hA = A >> 1, hB = B >> 1, pidx = (i >> 1) * hB + (j >> 1);
                d0 = 2 * mprev[pidx] + (i & 1);
                d1 = 2 * mprev[hA * hB + pidx] + (j & 1);
            }
            int r0, r1;
            float sc;
            near_match<float>(lv + (size_t)idx * cells, A, B, d0, d1, r0, r1, sc);
            if (k == 1) {
                a.match1[(size_t)n * 2 * cells + idx] = r0;
                a.match1[(size_t)n * 2 * cells + cells + idx] = r1;
            } else {
                mcur[idx] = r0;
                mcur[cells + idx] = r1;
            }
        }
        __syncthreads();
        int32_t* t = mprev; mprev = mcur; mcur = t;
    }
}

}  // namespace

// the level dm_upper_tail starts from (the last one dm_aggregate builds): the smallest k >= 2 whose
// per-tile level is <= 16 KiB
int dm_upper_tail_first_level(int t0, int t1, int levels) {
    int ks = 2;
    while (ks < levels - 1) {
        const size_t cells = (size_t)(t0 >> ks) * (t1 >> ks);
        if (cells * cells * sizeof(float) <= 16 * 1024) break;
        ++ks;
    }
    return ks;
}

bool dm_upper_tail_supported(int t0, int t1, int levels) {
    if (levels < 3 || levels > 16) return false;
    const int ks = dm_upper_tail_first_level(t0, t1, levels);
    size_t floats = 0;
    for (int k = ks + 1; k < levels; ++k) { const size_t c = (size_t)(t0 >> k) * (t1 >> k); floats += c * c; }
    const size_t bytes = floats * 4 + (size_t)4 * (t0 >> 2) * (t1 >> 2) * 4;
    return bytes <= 48 * 1024;
}

int dm_upper_tail(float* const* levels_dev, int n_tiles, int t0, int t1, int levels, int32_t* match1_dev, cudaStream_t stream) {
    DM_REQUIRE(dm_upper_tail_supported(t0, t1, levels), DM_ERR_UNSUPPORTED, "dm_upper_tail: unsupported grid (%d,%d), %d levels", t0, t1, levels);
    TailArgs a;
    for (int k = 0; k < 16; ++k) a.level[k] = k < levels ? levels_dev[k] : nullptr;
    a.t0 = t0; a.t1 = t1; a.L = levels; a.ks = dm_upper_tail_first_level(t0, t1, levels);
    a.match1 = match1_dev;
    size_t floats = 0;
    for (int k = a.ks + 1; k < levels; ++k) { const size_t c = (size_t)(t0 >> k) * (t1 >> k); floats += c * c; }
    const size_t smem = floats * 4 + (size_t)4 * (t0 >> 2) * (t1 >> 2) * 4;
    dm_launch_dep(DM_PDL_UPPER, dm_upper_tail_kernel, dim3((unsigned)n_tiles), dim3(256), smem, stream, a);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

// bad_matching.py:68-70: dis[i,j] = j - np.argmax(co_map[i, j, i, :]) -- the best column of a patch's OWN map row
// (a one-dimensional, same-row disparity search).  One warp per patch: first maximum, the first NaN wins
// (np.argmax).  Optionally also copies the row itself out (rows[n][i][j][:]), so that a host caller that
// indexes co_map[i, j, i, :] gets P rows of T1 values instead of the P x P map.
__global__ void __launch_bounds__(256)
dm_row_argmax_kernel(const float* __restrict__ co_map, long long n_patches, int T0, int T1,
                     int32_t* __restrict__ arg, float* __restrict__ rows) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_patches) return;
    const int P = T0 * T1;
    const int p = (int)(w % P), i = p / T1;
    const float* row = co_map + (size_t)w * P + (size_t)i * T1;
    float best = 0.f; int bi = 0x7fffffff; bool nan = false;      // bi = INT_MAX: nothing seen yet
    for (int x = lane; x < T1; x += 32) {
        const float v = row[x];
        if (rows) rows[(size_t)w * T1 + x] = v;
        const bool vn = v != v;
        if (bi == 0x7fffffff || (!nan && (vn || v > best))) { best = v; bi = x; nan = vn; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool on = __shfl_xor_sync(0xffffffffu, (int)nan, o) != 0;
        if (oi == 0x7fffffff) continue;
        // the other lane's candidate wins if it is a NaN at a smaller index (or mine is not a NaN),
        // or a larger value, or an equal value at a smaller index
        const bool take = bi == 0x7fffffff || (on && (!nan || oi < bi)) || (!on && !nan && (ob > best || (ob == best && oi < bi)));
        if (take) { best = ob; bi = oi; nan = on; }
    }
    if (lane == 0) arg[w] = bi;
}

template <typename T>
static int backtrack_launch(const void* level, long long n, int a, int b, int c, int d,
                            const int32_t* parent, int32_t* match, void* score, cudaStream_t st) {
    const long long total = n * a * b;
    dm_backtrack_kernel<T><<<dm_div_up(total, 256), 256, 0, st>>>((const T*)level, total, a, b, c, d, parent, match, (T*)score);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_backtrack_top(const void* top_dev, int is_f64, int n, int a, int b,
                                int32_t* match_dev, void* score_dev, void* stream) {
    DM_REQUIRE(n > 0 && a > 0 && b > 0, DM_ERR_INVALID, "dm_backtrack_top: bad shape");
    return is_f64 ? backtrack_launch<double>(top_dev, n, a, b, a, b, nullptr, match_dev, score_dev, (cudaStream_t)stream)
                  : backtrack_launch<float>(top_dev, n, a, b, a, b, nullptr, match_dev, score_dev, (cudaStream_t)stream);
}

extern "C" int dm_backtrack_level(const void* level_dev, int is_f64, int n, int a, int b, int c, int d,
                                  const int32_t* parent_match_dev, int32_t* match_dev, void* score_dev, void* stream) {
    DM_REQUIRE(n > 0 && a >= 2 && b >= 2 && !(a & 1) && !(b & 1) && c > 0 && d > 0, DM_ERR_INVALID, "dm_backtrack_level: bad shape");
    DM_REQUIRE(parent_match_dev != nullptr, DM_ERR_INVALID, "dm_backtrack_level: parent matches missing");
    return is_f64 ? backtrack_launch<double>(level_dev, n, a, b, c, d, parent_match_dev, match_dev, score_dev, (cudaStream_t)stream)
                  : backtrack_launch<float>(level_dev, n, a, b, c, d, parent_match_dev, match_dev, score_dev, (cudaStream_t)stream);
}

extern "C" int dm_match_filter(const int32_t* match_in_dev, int n, int h, int w, int window, int mode,
                               int32_t* match_out_dev, void* stream) {
    DM_REQUIRE(n > 0 && h > 0 && w > 0 && window >= 1, DM_ERR_INVALID, "dm_match_filter: bad shape");
    DM_REQUIRE(mode == DM_FILTER_MEDIAN || mode == DM_FILTER_AVERAGE, DM_ERR_INVALID, "dm_match_filter: invalid mode %d", mode);
    DM_REQUIRE(match_in_dev != match_out_dev, DM_ERR_INVALID, "dm_match_filter: in-place filtering is not possible");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)n * h * w;
    if (!(h >= window && w >= window)) {        // misc/Matching.py:230: smaller maps pass through
        DM_CUDA_CHECK(cudaMemcpyAsync(match_out_dev, match_in_dev, (size_t)total * 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        return DM_OK;
    }
    DM_REQUIRE(h == w, DM_ERR_UNSUPPORTED, "dm_match_filter: Matching._filter sizes its snapshot (shape[1], shape[1]) and is undefined on non-square maps (%d x %d)", h, w);
    const int e = (window - 1) / 2;
    dm_match_filter_kernel<<<dm_div_up(total, 256), 256, 0, st>>>(match_in_dev, total, h, w, e, mode, match_out_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_match_map(const void* level0_dev, int is_f64, int n, int t0, int t1,
                            const int32_t* match_dev, const void* score_dev, int sub_pix,
                            double* map_dev, void* stream) {
    DM_REQUIRE(n > 0 && t0 > 0 && t1 > 0, DM_ERR_INVALID, "dm_match_map: bad shape");
    const long long total = (long long)n * t0 * t1;
    cudaStream_t st = (cudaStream_t)stream;
    DM_REQUIRE(is_f64 >= 0 && is_f64 <= 2, DM_ERR_INVALID, "dm_match_map: is_f64 %d", is_f64);
    if (is_f64 == 1)
        dm_match_map_kernel<double, false><<<dm_div_up(total, 256), 256, 0, st>>>((const double*)level0_dev, total, t0, t1, match_dev, (const double*)score_dev, sub_pix, map_dev);
    else if (is_f64 == 2)
        dm_match_map_kernel<float, true><<<dm_div_up(total, 256), 256, 0, st>>>((const float*)level0_dev, total, t0, t1, match_dev, (const float*)score_dev, sub_pix, map_dev);
    else
        dm_match_map_kernel<float, false><<<dm_div_up(total, 256), 256, 0, st>>>((const float*)level0_dev, total, t0, t1, match_dev, (const float*)score_dev, sub_pix, map_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_cal_map(const double* map_dev, int t0, int t1, int mode, double* out_dev, void* stream) {
    DM_REQUIRE(mode >= 0 && mode <= 2, DM_ERR_INVALID, "dm_cal_map: invalid mode %d", mode);
    dm_cal_map_kernel<<<dm_div_up((long long)t0 * t1, 256), 256, 0, (cudaStream_t)stream>>>(map_dev, t0, t1, mode, out_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_sub_pix_cal(const double* arr_dev, const double* co_map_dev, int s0, int s1,
                              int direction, double ratio, double* out_dev, void* stream) {
    DM_REQUIRE(s0 > 0 && s1 > 0 && (direction == 0 || direction == 1), DM_ERR_INVALID, "dm_sub_pix_cal: bad arguments");
    dm_sub_pix_cal_kernel<<<dm_div_up((long long)s0 * s1, 256), 256, 0, (cudaStream_t)stream>>>(arr_dev, co_map_dev, s0, s1, direction, ratio, out_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_row_argmax(const float* co_map_dev, int n, int t0, int t1, int32_t* arg_dev, float* rows_dev, void* stream) {
    DM_REQUIRE(co_map_dev && arg_dev && n > 0 && t0 > 0 && t1 > 0, DM_ERR_INVALID, "dm_row_argmax: bad arguments");
    const long long n_patches = (long long)n * t0 * t1;
    dm_row_argmax_kernel<<<dm_div_up(n_patches, 8), 256, 0, (cudaStream_t)stream>>>(co_map_dev, n_patches, t0, t1, arg_dev, rows_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}
