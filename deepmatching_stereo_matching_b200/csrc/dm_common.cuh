// Shared device helpers and host-side error plumbing for libdmstereo (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dmstereo.h"

#define DM_LAM 1.4f                 // misc/Correlation_map.py:41
#define DM_NEAR_ZERO_F 0.0001f      // misc/Matching.py:74
#define DM_NEAR_ZERO_D 0.0001

void dm_set_error(const char* fmt, ...);

#define DM_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            dm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),      \
                         __FILE__, __LINE__);                                            \
            return DM_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define DM_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            dm_set_error(__VA_ARGS__);                                                   \
            return (code);                                                               \
        }                                                                                \
    } while (0)

#define DM_LAUNCH_CHECK() DM_CUDA_CHECK(cudaGetLastError())

// ---- per-patch statistics written by the descriptor kernel -------------------------
// x = S'  (residual sum of the mean-centred window, |S'| <= K/2)
// y = inv (1/sqrt(sum a'^2 - S'^2/K); 0 exactly when the window is flat)
// z = S'/K
// w = the rounded mean the window was centred on (integer 0..255)
typedef float4 dm_stat;

// NaN-propagating max/min: torch.nn.MaxPool2d and numpy max/min keep NaN
// (misc/Correlation_map.py:103, misc/Feature_value.py:34-35); fmaxf would drop it.
__device__ __forceinline__ float dm_max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float dm_min_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// ---- the one ZNCC formula every kernel shares ---------------------------------------
// dot = sum a1'*a2' (exact integer in fp32), returns z = (dot - S1'*S2'/K) * inv2.
// The row factor inv1 (>0) and the [-1,1] clamp are monotone, so kernels that max-pool
// apply dm_zncc_finish after pooling and still get bit-identical values.
__device__ __forceinline__ float dm_zncc_partial(float dot, float s1, float s2k, float inv2) {
    return __fmul_rn(__fmaf_rn(-s1, s2k, dot), inv2);
}
__device__ __forceinline__ float dm_zncc_finish(float z, float inv1, bool flat1, bool normed) {
    if (!normed) return z;                       // TM_CCOEFF: numerator only (inv2 == 1)
    if (flat1) return 1.0f;                      // OpenCV: flat template -> map of ones
    float r = __fmul_rn(z, inv1);
    return fminf(fmaxf(r, -1.0f), 1.0f);
}

// (x - min) / (max - min)  (misc/Feature_value.py:36) evaluated as (x - min) * rinv with
// rinv = 1/(max - min) rounded once per slice: within 1 ulp of the float32 division, and
// 0 * inf = NaN keeps the reference's NaN slice for a flat patch (max == min).
__device__ __forceinline__ float dm_range_inv(float mn, float mx) { return __frcp_rn(__fsub_rn(mx, mn)); }
__device__ __forceinline__ float dm_normalize(float x, float mn, float mx, float rinv) {
    const float r = __fmul_rn(__fsub_rn(x, mn), rinv);
    return x == mx ? (mx > mn ? 1.0f : r) : r;        // the slice maximum maps to exactly 1, as with a division
}
// x ** 1.4 (misc/Correlation_map.py:159) for x in [0,1] as ex2(1.4 * lg2 x) on the
// special-function unit: 0 -> 0 and 1 -> 1 exactly, NaN propagates, absolute error < 2e-7
// (relative < 3e-6 down to x = 1e-3).  libdevice powf costs ~10x more and made the pyramid
// kernels ALU-bound instead of HBM-bound.
__device__ __forceinline__ float dm_rectify(float x) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    l = __fmul_rn(DM_LAM, l);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
    return r;
}

// misc/Matching.py:165-175: the parabola through (-1, rm), (0, r0), (1, r1).  The reference holds level
// 0 in float64, so the differences are formed in float64 here as well: with a flat peak
// (r1 + rm - 2 r0 -> 0) the float32 subtraction would lose the few digits that are left.
__device__ __forceinline__ double dm_parabola_shift(float r0, float r1, float rm) {
    const double a = (double)r1, b = (double)rm, c = (double)r0;
    return -(a - b) / (2.0 * (a + b - 2.0 * c));
}

// numpy indexing of an axis of length n: -n <= v < n is accepted, negative values wrap
__device__ __forceinline__ bool dm_np_index_ok(int v, int n) { return v >= -n && v < n; }
__device__ __forceinline__ int dm_np_wrap(int v, int n) { return v < 0 ? v + n : v; }

__device__ __forceinline__ int dm_round_mean(int sum, int k) {
    // nearest integer to sum/k for sum >= 0 (pixels are unsigned)
    return (2 * sum + k) / (2 * k);
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------
// The launches of a chunk form one dependent chain on one stream.  A kernel launched through dm_launch_dep
// carries cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may become resident -- and run whatever
// precedes dm_pdl_wait(): barrier init, TMEM allocation, tensor-map prefetch -- while the last wave of the kernel in
// front of it drains.  dm_pdl_wait() returns once that kernel has COMPLETED and its writes are visible, so the data
// dependency is exactly the stream order; every kernel of the chain calls it first thing, before any early
// return, which keeps completion transitive (K_c waits for K_b, whose threads all waited for K_a).  Launched
// without the attribute both instructions are no-ops.  dm_pdl_launch_dependents() early in a kernel only lets the
// next grid be scheduled once every CTA of this one is resident or done; it takes nothing away from this grid.
__device__ __forceinline__ void dm_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void dm_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
enum { DM_PDL_DESC = 1, DM_PDL_CORR = 2, DM_PDL_FIRST = 4, DM_PDL_UPPER = 8, DM_PDL_FINAL = 16 };      // which launches carry the attribute
bool dm_pdl_enabled(int which); // capi.cu: DM_PDL = bit mask over the launches above
void dm_pdl_suppress(bool off); // this thread launches without the attribute until told otherwise (graph-capture fallback)

template <typename... KArgs, typename... Args>
static inline cudaError_t dm_launch_dep(int which, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    if (dm_pdl_enabled(which)) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

static inline int dm_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Exact unsigned 32-bit division by a runtime constant (Granlund & Montgomery, round-up method):
// q = (t + ((x - t) >> s1)) >> s2 with t = umulhi(m, x); valid for every 32-bit x.  The constants
// are made on the host once per launch; a runtime-divisor `/` costs ~20 instructions (32-bit) to
// ~80 (64-bit), and index arithmetic of that kind was 15-30 % of two kernels' instructions.
struct dm_fastdiv { uint32_t m, s1, s2; };
static inline dm_fastdiv dm_make_fastdiv(uint32_t d) {
    dm_fastdiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                          // ceil(log2 d)
    f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l > 0 ? l - 1 : 0;
    return f;
}
__device__ __forceinline__ uint32_t dm_fd_div(uint32_t x, const dm_fastdiv& f) {
    const uint32_t t = __umulhi(f.m, x);
    return (t + ((x - t) >> f.s1)) >> f.s2;
}
