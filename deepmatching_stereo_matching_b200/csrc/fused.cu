// Fused scene solver: the level-0 correlation map never reaches HBM.
//
//   descriptors -> tcgen05 correlation with pooled epilogue (correlation_umma.cu, MODE_POOL)
//   -> first aggregation from the pooled raw map (this file) -> generic aggregation for
//   the upper levels (pyramid.cu) -> backtracking down to level 1 (backtrack.cu) -> final
//   level: the 3x3 window of level-0 values around each predicted match and the four
//   parabola neighbours are recomputed from the two images (exact integer dot products,
//   the same ZNCC / min-max / **1.4 formula as everywhere else), then the disparity planes
//   are written straight into the mosaic.
//
// Replaces, for a batch of tiles, Correlation_map.__call__ (misc/Correlation_map.py:161-173),
// Matching.__call__ (misc/Matching.py:211-222), Calc_difference.cal_map
// (misc/Calc_difference.py:25-49) and the paste of misc/image_cut_solver.py:165-175.
// HBM traffic per tile drops from ~2.1 * 4 P^2 bytes (write raw, read+write level 0, read
// level 0) to ~0.6 * 4 P^2.
#include <stdlib.h>

#include "dm_common.cuh"
#include "dm_internal.h"

namespace {

struct FusedBuffers {
    int32_t* origin; int32_t* tinfo; void* desc1; void* desc2; float* stat1; float* stat2;
    float* pooled; float* rowmin; float* rowmax;
    float* level[16];           // level[0] unused
    int32_t* match[2]; float* score[2];
};

struct Carve {
    char* base; size_t off;
    explicit Carve(char* b) : base(b), off(0) {}
    template <typename T> T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? (T*)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

size_t carve(char* base, int nt, int t0, int t1, int kpad, int levels, FusedBuffers& fb) {
    Carve c(base);
    const size_t P = (size_t)t0 * t1;
    fb.origin = c.take<int32_t>((size_t)nt * 2);
    fb.tinfo = c.take<int32_t>((size_t)nt * 4);
    fb.desc1 = c.take<__nv_bfloat16>((size_t)nt * P * kpad);
    fb.desc2 = c.take<__nv_bfloat16>((size_t)nt * P * kpad);
    fb.stat1 = c.take<float>((size_t)nt * P * DM_STAT_FLOATS);
    fb.stat2 = c.take<float>((size_t)nt * P * DM_STAT_FLOATS);
    fb.pooled = c.take<float>((size_t)nt * P * (P / 4));
    fb.rowmin = c.take<float>((size_t)nt * P * 4);      // four partial min / max per patch (column quarters; the 2-way kernels duplicate)
    fb.rowmax = c.take<float>((size_t)nt * P * 4);
    size_t a = t0 >> 1, b = t1 >> 1;
    fb.level[0] = nullptr;
    for (int k = 1; k < levels; ++k) {
        fb.level[k] = c.take<float>((size_t)nt * a * b * a * b);
        a >>= 1; b >>= 1;
    }
    for (int s = 0; s < 2; ++s) {
        fb.match[s] = c.take<int32_t>((size_t)nt * 2 * (P / 4));
        fb.score[s] = c.take<float>((size_t)nt * (P / 4));
    }
    return (c.off + 255) & ~(size_t)255;
}

// tile g of a (batch of) scene(s) stacked along rows: origin = (scene*S0 + s0*gi, s1*gj)
// tinfo = (tile row gi, tile column gj, scene, 0): the final-level kernel needs them per quad and
// would otherwise redo these divisions in every warp
__global__ void dm_tile_origin_kernel2(int32_t* origin, int32_t* tinfo, int n, int first_tile, int len0, int len1, int s0, int s1, int scene_h) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int g = first_tile + t, tps = len0 * len1;
    const int sc = g / tps, r = g - sc * tps;
    const int gi = r / len1, gj = r - gi * len1;
    origin[2 * t] = sc * scene_h + s0 * gi;
    origin[2 * t + 1] = s1 * gj;
    reinterpret_cast<int4*>(tinfo)[t] = make_int4(gi, gj, sc, 0);
}

// ---------------------------------------------------------------------------------------
// First aggregation from the pooled raw ZNCC: per child min-max + **1.4 (monotone, so they
// commute with the max-pool the GEMM epilogue already did), average of the four children,
// **1.4.  (misc/Feature_value.py:36, misc/Correlation_map.py:109-128,158-159)
// pooled [n][t0*t1][t0/2][t1/2] -> level 1 [n][t0/2][t1/2][t0/2][t1/2]; 4 outputs / thread.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dm_aggregate_first_kernel(const float* __restrict__ pooled, const float* __restrict__ rowmin,
                          const float* __restrict__ rowmax, int t0, int t1, float* __restrict__ out) {
    dm_pdl_wait();                  // the pooled map and the row minima / maxima come from the correlation kernel
    dm_pdl_launch_dependents();
    // blockIdx.x = parent (n, I, J) flattened; threads cover the P/16 float4 of its map
    const int P = t0 * t1, Q4 = P >> 4;
    const int hA = t0 >> 1, hB = t1 >> 1;
    const unsigned par = blockIdx.x;
    const int J = par % hB;
    const unsigned t = par / hB;
    const int I = t % hA;
    const size_t n = t / hA;
    const float4* src[4];
    float mn[4], mx[4], rinv[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const size_t p = n * P + (size_t)(2 * I + (ch >> 1)) * t1 + (2 * J + (ch & 1));
        const float4 pmn = __ldg(reinterpret_cast<const float4*>(rowmin) + p), pmx = __ldg(reinterpret_cast<const float4*>(rowmax) + p);
        mn[ch] = dm_min_nan(dm_min_nan(pmn.x, pmn.y), dm_min_nan(pmn.z, pmn.w));
        mx[ch] = dm_max_nan(dm_max_nan(pmx.x, pmx.y), dm_max_nan(pmx.z, pmx.w));
        rinv[ch] = dm_range_inv(mn[ch], mx[ch]);
        src[ch] = reinterpret_cast<const float4*>(pooled + p * (size_t)(P / 4));
    }
    float4* dst = reinterpret_cast<float4*>(out) + (size_t)par * Q4;
    for (int m4 = threadIdx.x; m4 < Q4; m4 += blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) v[ch] = __ldg(src[ch] + m4);
        float4 sum;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float4 r;
            r.x = dm_rectify(dm_normalize(v[ch].x, mn[ch], mx[ch], rinv[ch]));
            r.y = dm_rectify(dm_normalize(v[ch].y, mn[ch], mx[ch], rinv[ch]));
            r.z = dm_rectify(dm_normalize(v[ch].z, mn[ch], mx[ch], rinv[ch]));
            r.w = dm_rectify(dm_normalize(v[ch].w, mn[ch], mx[ch], rinv[ch]));
            if (ch == 0) sum = r;
            else { sum.x = __fadd_rn(sum.x, r.x); sum.y = __fadd_rn(sum.y, r.y); sum.z = __fadd_rn(sum.z, r.z); sum.w = __fadd_rn(sum.w, r.w); }
        }
        float4 o;
        o.x = dm_rectify(__fmul_rn(sum.x, 0.25f)); o.y = dm_rectify(__fmul_rn(sum.y, 0.25f));
        o.z = dm_rectify(__fmul_rn(sum.z, 0.25f)); o.w = dm_rectify(__fmul_rn(sum.w, 0.25f));
        dst[m4] = o;
    }
}

// Small maps (P/16 <= 64 float4 per parent map, image_size <= 32): one WARP per parent, warps striding over the
// parents of a persistent grid.  With one 64-thread CTA per parent a 64 x 512^2 batch is 3.2 million CTAs of ~400
// instructions each, and the rate at which CTAs can be handed out -- not HBM, not the instruction count -- bounds
// the kernel (four parents per CTA: -5 %; cooperative statistics alone: nothing).  A lane owns TRIPS float4
// columns of the parent's map; all its loads (TRIPS x 4 map loads + the statistics of child lane & 3) are issued
// before any arithmetic.  Same operations in the same order as dm_aggregate_first_kernel: bit-identical level 1.
template <int TRIPS>
__global__ void __launch_bounds__(256, 4)
dm_aggregate_first_small_kernel(const float* __restrict__ pooled, const float* __restrict__ rowmin,
                                const float* __restrict__ rowmax, int t0, int t1, unsigned n_parents,
                                dm_fastdiv fd_hb, dm_fastdiv fd_ha, float* __restrict__ out) {
    dm_pdl_wait();
    dm_pdl_launch_dependents();
    const int P = t0 * t1, Q4 = P >> 4;
    const int hA = t0 >> 1, hB = t1 >> 1;
    const int lane = threadIdx.x & 31, c = lane & 3;
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    for (unsigned par = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); par < n_parents; par += warps) {
        const unsigned t = dm_fd_div(par, fd_hb);
        const int J = (int)(par - t * (unsigned)hB);
        const unsigned nn = dm_fd_div(t, fd_ha);
        const int I = (int)(t - nn * (unsigned)hA);
        const size_t p00 = (size_t)nn * P + (size_t)(2 * I) * t1 + 2 * J;
        const float4* src[4];
        src[0] = reinterpret_cast<const float4*>(pooled) + p00 * (size_t)Q4;
        src[1] = src[0] + Q4;
        src[2] = src[0] + (size_t)t1 * (size_t)Q4;
        src[3] = src[2] + Q4;
        float4 v[TRIPS][4];
#pragma unroll
        for (int tr = 0; tr < TRIPS; ++tr) {
            const int m4 = lane + 32 * tr;
            if (m4 < Q4) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) v[tr][ch] = __ldg(src[ch] + m4);
            }
        }
        const size_t p = p00 + (size_t)(c >> 1) * t1 + (c & 1);
        const float4 pmn = __ldg(reinterpret_cast<const float4*>(rowmin) + p), pmx = __ldg(reinterpret_cast<const float4*>(rowmax) + p);
        const float cmn = dm_min_nan(dm_min_nan(pmn.x, pmn.y), dm_min_nan(pmn.z, pmn.w));
        const float cmx = dm_max_nan(dm_max_nan(pmx.x, pmx.y), dm_max_nan(pmx.z, pmx.w));
        const float cinv = dm_range_inv(cmn, cmx);
        float mn[4], mx[4], rinv[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            mn[ch] = __shfl_sync(0xffffffffu, cmn, ch);
            mx[ch] = __shfl_sync(0xffffffffu, cmx, ch);
            rinv[ch] = __shfl_sync(0xffffffffu, cinv, ch);
        }
        float4* dst = reinterpret_cast<float4*>(out) + (size_t)par * Q4;
#pragma unroll
        for (int tr = 0; tr < TRIPS; ++tr) {
            const int m4 = lane + 32 * tr;
            if (m4 < Q4) {
                float4 sum;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    float4 r;
                    r.x = dm_rectify(dm_normalize(v[tr][ch].x, mn[ch], mx[ch], rinv[ch]));
                    r.y = dm_rectify(dm_normalize(v[tr][ch].y, mn[ch], mx[ch], rinv[ch]));
                    r.z = dm_rectify(dm_normalize(v[tr][ch].z, mn[ch], mx[ch], rinv[ch]));
                    r.w = dm_rectify(dm_normalize(v[tr][ch].w, mn[ch], mx[ch], rinv[ch]));
                    if (ch == 0) sum = r;
                    else { sum.x = __fadd_rn(sum.x, r.x); sum.y = __fadd_rn(sum.y, r.y); sum.z = __fadd_rn(sum.z, r.z); sum.w = __fadd_rn(sum.w, r.w); }
                }
                float4 o;
                o.x = dm_rectify(__fmul_rn(sum.x, 0.25f)); o.y = dm_rectify(__fmul_rn(sum.y, 0.25f));
                o.z = dm_rectify(__fmul_rn(sum.z, 0.25f)); o.w = dm_rectify(__fmul_rn(sum.w, 0.25f));
                dst[m4] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Final level: arguments.
// ---------------------------------------------------------------------------------------
struct FinalArgs {
    const uint8_t* img1; const uint8_t* img2; int pitch;
    const int32_t* origin; const int32_t* tinfo; const dm_stat* stat1; const dm_stat* stat2;
    dm_fastdiv fd_pq, fd_hb, fd_s0, fd_s1;     // divisors PQ = (t0/2)(t1/2), t1/2, s0, s1
    dm_fastdiv fd_p, fd_t1;                    // divisors P = t0 t1, t1 (patch-per-thread kernel)
    const float* rowmin; const float* rowmax;
    const int32_t* parent;      // level-1 matches [n][2][t0/2][t1/2]
    int t0, t1, ws, normed, sub_pix, scene_h;
    int n_modes, modes[4];
    int s0, s1, len0, len1, out_h, out_w, first_tile;
    int tile0;                  // first tile of this launch (the final stage may run in bands of whole tiles)
    double* d_map; double* out_map;
};

// ---------------------------------------------------------------------------------------
// Final level, one warp per PARENT: its four children (2I+ci, 2J+cj) share the parent's
// match, so their candidate windows lie in one (ws+5)^2 region of image 2.  The warp stages
// that region once (two copies, the second shifted by one byte so that both cj = 0 and
// cj = 1 see compile-time funnel shifts); 8 lanes serve one child, lane l owns patch rows
// l and l+8 as 4 packed words each.  Reductions, candidate evaluation, argmax, parabola
// fit and the plane writes are shared by the four children.
// ---------------------------------------------------------------------------------------
constexpr int QRS = 32;                 // region row stride in bytes (16-byte aligned rows)
constexpr int QROWS = 20;               // ws + 5 <= 20

constexpr int FQ_WARPS = 4;             // quads (warps) per CTA: 0.371 -> 0.359 ms against 8

template <int WS>
__global__ void __launch_bounds__(32 * FQ_WARPS)
dm_final_quad_kernel(const FinalArgs a, long long n_quads) {
    constexpr int K = WS * WS;
    constexpr int RWQ = WS + 5;
    __shared__ __align__(16) uint8_t region_all[FQ_WARPS][QROWS * QRS];
    __shared__ __align__(16) uint8_t patch_all[FQ_WARPS][16 * QRS];         // the quad's (ws+1)^2 block of image 1
    const int lane = threadIdx.x & 31;
    uint8_t* reg0 = region_all[threadIdx.x >> 5];
    // quad index relative to this launch (a launch starts at a tile boundary and never holds 2^32 quads)
    const long long wl = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wl >= n_quads) return;
    const uint32_t wq = (uint32_t)wl;
    const int T0 = a.t0, T1 = a.t1, P = T0 * T1;
    const int hA = T0 >> 1, hB = T1 >> 1, PQ = hA * hB;
    const uint32_t nrel = dm_fd_div(wq, a.fd_pq);
    const int n = a.tile0 + (int)nrel;                   // tile inside the chunk
    const int pq = (int)(wq - nrel * (uint32_t)PQ);
    const int I = (int)dm_fd_div((uint32_t)pq, a.fd_hb), J = pq - I * hB;
    const int4 ti = __ldg(reinterpret_cast<const int4*>(a.tinfo) + n);      // (gi, gj, scene, -)
    {
        // misc/image_cut_solver.py:165-175: a pixel belongs to the covering tile with the largest
        // index.  A quad whose two patch rows (or columns) lie at or beyond the stride is pasted
        // over by the next tile: nothing of it survives, so the warp stops here (12 % of the
        // quads at image_size 64 / stride 60).
        const int gi = ti.x, gj = ti.y;
        if ((gi < a.len0 - 1 && 2 * I >= a.s0) || (gj < a.len1 - 1 && 2 * J >= a.s1)) return;
    }
    const int oy = a.origin[2 * n], ox = a.origin[2 * n + 1];
    const bool normed = a.normed != 0;
    const int c = lane >> 3, l = lane & 7, ci = c >> 1, cj = c & 1;
    const unsigned gbase = lane & 24;                    // first lane of this child's group
    const int i = 2 * I + ci, j = 2 * J + cj, p = i * T1 + j;

    // misc/Matching.py:116-124: p_dot = 2 * parent match + o
    const size_t pb = (size_t)n * 2 * PQ;
    const int pm0 = a.parent[pb + pq], pm1 = a.parent[pb + PQ + pq];
    const int d0 = 2 * pm0 + ci, d1 = 2 * pm1 + cj;

    const dm_stat s1 = a.stat1[(size_t)n * P + p];
    const bool flat1 = (s1.y == 0.0f);
    const float4 pmn = reinterpret_cast<const float4*>(a.rowmin)[(size_t)n * P + p], pmx = reinterpret_cast<const float4*>(a.rowmax)[(size_t)n * P + p];
    const float mn = dm_min_nan(dm_min_nan(pmn.x, pmn.y), dm_min_nan(pmn.z, pmn.w));
    const float mx = dm_max_nan(dm_max_nan(pmx.x, pmx.y), dm_max_nan(pmx.z, pmx.w)), rinv = dm_range_inv(mn, mx);
    const dm_stat* st2 = a.stat2 + (size_t)n * P;
    const int m1 = (int)s1.w, S1 = (int)s1.x;

    // Small windows (ws <= 7) are instruction-bound here: a lane then owns one ROW, fetches its bytes
    // as (at most four) aligned 32-bit words -- only words that hold a needed byte -- and stores them
    // with one or two 16-byte stores (measured on 64 x 512^2, ws 5: 3.97 -> 3.77 ms).  For large
    // windows the row-per-lane loads touch ws + 5 cache lines per request and the kernel, L1-bound
    // there, got slower (ws 15: 0.42 -> 0.48 ms), so they keep the coalesced byte loads.
    constexpr bool ROW_STAGING = WS <= 7;
    auto load_row = [&](const uint8_t* src, int nbytes, bool on, uint32_t (&c0)[4]) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
        const int off = (int)(addr & 3);
        uint32_t w[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (on && 4 * k < off + nbytes) ? __ldg(ap + k) : 0u;
        w[4] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int nvk = nbytes - 4 * k;             // valid bytes of word k
            const uint32_t m = nvk >= 4 ? 0xffffffffu : (nvk <= 0 ? 0u : ((1u << (8 * nvk)) - 1u));
            c0[k] = __funnelshift_r(w[k], w[k + 1], 8 * off) & m;
        }
    };
    // ---- stage the shared region: scene rows oy+2*pm0-2 .. +RWQ, cols ox+2*pm1-2 .. +RWQ
    {
        const int gy0 = oy + 2 * pm0 - 2, gx0 = ox + 2 * pm1 - 2;
        uint8_t vals[RWQ];
        const bool inside = gy0 >= 0 && gy0 + RWQ <= a.scene_h && gx0 >= 0 && gx0 + RWQ <= a.pitch;    // warp-uniform
        if (ROW_STAGING && inside) {
            uint32_t c0[4];                              // RWQ <= 12 bytes
            const bool on = lane < RWQ;
            load_row(a.img2 + (size_t)(gy0 + (on ? lane : 0)) * a.pitch + gx0, RWQ, on, c0);
            if (on) {
                uint4* rp = reinterpret_cast<uint4*>(reg0 + lane * QRS);
                rp[0] = make_uint4(c0[0], c0[1], c0[2], c0[3]);
                rp[1] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
        if (inside) {
            const uint8_t* src = a.img2 + (size_t)gy0 * a.pitch + gx0 + (lane < RWQ ? lane : 0);
#pragma unroll
            for (int ry = 0; ry < RWQ; ++ry) { vals[ry] = __ldg(src); src += a.pitch; }
        } else {
            const int gx = gx0 + lane;
            const bool colok = lane < RWQ && gx >= 0 && gx < a.pitch;
#pragma unroll
            for (int ry = 0; ry < RWQ; ++ry) {
                const int gy = gy0 + ry;
                vals[ry] = (colok && gy >= 0 && gy < a.scene_h) ? __ldg(a.img2 + (size_t)gy * a.pitch + gx) : (uint8_t)0;
            }
        }
#pragma unroll
        for (int ry = 0; ry < RWQ; ++ry) {
            const uint8_t v = (lane < RWQ) ? vals[ry] : (uint8_t)0;
            reg0[ry * QRS + lane] = v;                   // column x at byte x, zero beyond the region
        }
        }
    }
    // ---- the four patches of the quad overlap in a (ws+1)^2 block of image 1: stage it with
    // row-coalesced loads (column x at byte x; the children with cj = 1 shift by one byte when they read)
    {
        uint8_t* pat0 = patch_all[threadIdx.x >> 5];
        if (ROW_STAGING) {
            uint32_t c0[4];                              // WS + 1 <= 8 bytes
            const bool on = lane <= WS;
            load_row(a.img1 + (size_t)(oy + 2 * I + (on ? lane : 0)) * a.pitch + ox + 2 * J, WS + 1, on, c0);
            if (on) {
                uint4* pp = reinterpret_cast<uint4*>(pat0 + lane * QRS);
                pp[0] = make_uint4(c0[0], c0[1], c0[2], c0[3]);
                pp[1] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
        const uint8_t* src = a.img1 + (size_t)(oy + 2 * I) * a.pitch + ox + 2 * J + (lane <= WS ? lane : 0);
        uint8_t pv[WS + 1];
#pragma unroll
        for (int ry = 0; ry <= WS; ++ry) { pv[ry] = __ldg(src); src += a.pitch; }
#pragma unroll
        for (int ry = 0; ry <= WS; ++ry) {
            const uint8_t v = (lane <= WS) ? pv[ry] : (uint8_t)0;
            pat0[ry * QRS + lane] = v;
        }
        }
    }
    __syncwarp();
    // ---- this lane's patch rows ky = l and l + 8, 16 bytes each (zero beyond the window)
    const int psh = 8 * cj;                              // this child's columns start at byte cj of a staged row
    uint32_t aw[2][4];
    {
        const uint8_t* patc = patch_all[threadIdx.x >> 5];
        // bytes >= WS of a row belong to the neighbouring child: masked out
        constexpr uint32_t M0 = WS >= 4 ? 0xffffffffu : ((1u << (8 * WS)) - 1u);
        constexpr uint32_t M1 = WS >= 8 ? 0xffffffffu : (WS <= 4 ? 0u : ((1u << (8 * (WS - 4))) - 1u));
        constexpr uint32_t M2 = WS >= 12 ? 0xffffffffu : (WS <= 8 ? 0u : ((1u << (8 * (WS - 8))) - 1u));
        constexpr uint32_t M3 = WS >= 16 ? 0xffffffffu : (WS <= 12 ? 0u : ((1u << (8 * (WS - 12))) - 1u));
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ky = l + 8 * r;
            const bool live = ky < WS;
            const uint4 w = *reinterpret_cast<const uint4*>(patc + (ci + (live ? ky : 0)) * QRS);
            const uint32_t w4 = *reinterpret_cast<const uint32_t*>(patc + (ci + (live ? ky : 0)) * QRS + 16);
            aw[r][0] = live ? (__funnelshift_r(w.x, w.y, psh) & M0) : 0u; aw[r][1] = live ? (__funnelshift_r(w.y, w.z, psh) & M1) : 0u;
            aw[r][2] = live ? (__funnelshift_r(w.z, w.w, psh) & M2) : 0u; aw[r][3] = live ? (__funnelshift_r(w.w, w4, psh) & M3) : 0u;
        }
    }
    __syncwarp();
    const uint8_t* regc = reg0;                          // region column x at byte x; this child's windows start cj bytes in

    // min-maxed (not yet rectified) co_map value of position (qy,qx) from sum a*b.  Candidates are
    // compared on these values: x -> x**1.4 is strictly increasing, and the reference compares its
    // float64 powers, which stay distinct where the float32 ex2/lg2 powers of two neighbouring
    // floats may coincide.
    auto value_at = [&](int sum_ab, int qy, int qx) -> float {
        const dm_stat sq = st2[qy * T1 + qx];
        const int m2 = (int)sq.w, S2 = (int)sq.x;
        const int dot = sum_ab - m2 * S1 - m1 * S2 - K * m1 * m2;
        const float z = dm_zncc_partial((float)dot, s1.x, sq.z, normed ? sq.y : 1.0f);
        // mn / mx come from the tensor-core kernel, whose correction terms are accumulated with a few ulp of
        // rounding; the value recomputed here is exact and may undercut mn by that much -- it is the slice
        // minimum then, i.e. 0 (a negative base would turn x ** 1.4 into NaN).  NaN stays NaN.
        const float r = dm_normalize(dm_zncc_finish(z, s1.y, flat1, normed), mn, mx, rinv);
        return r < 0.0f ? 0.0f : r;
    };
    // 16 bytes of a against region row `row`, window starting at byte `bo` (dynamic)
    auto row_dot = [&](const uint32_t (&aq)[4], int row, int bo) -> uint32_t {
        const uint32_t* rr = reinterpret_cast<const uint32_t*>(regc + row * QRS) + ((bo + cj) >> 2);
        const int sh = ((bo + cj) & 3) * 8;
        const uint32_t w0 = rr[0], w1 = rr[1], w2 = rr[2], w3 = rr[3], w4 = rr[4];
        uint32_t acc = __dp4a(aq[0], __funnelshift_r(w0, w1, sh), 0u);
        acc = __dp4a(aq[1], __funnelshift_r(w1, w2, sh), acc);
        acc = __dp4a(aq[2], __funnelshift_r(w2, w3, sh), acc);
        return __dp4a(aq[3], __funnelshift_r(w3, w4, sh), acc);
    };

    // ---- misc/Matching.py:58-78 on the recomputed 3x3 window (zero padding outside)
    // candidate (dy,dx) of this child starts at region row ci + dy + 2, byte dx + 2 of its view
    uint32_t acc[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) acc[s] = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int ky = l + 8 * r;
        if (ky < WS) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const uint4 lo = *reinterpret_cast<const uint4*>(regc + (ky + ci + dy + 1) * QRS);
                const uint32_t w4 = *reinterpret_cast<const uint32_t*>(regc + (ky + ci + dy + 1) * QRS + 16);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int sh = 8 * (dx + 1) + psh;       // 8 .. 32: the clamping funnel shift returns the high word at 32
                    uint32_t t = __dp4a(aw[r][0], __funnelshift_rc(lo.x, lo.y, sh), acc[dy * 3 + dx]);
                    t = __dp4a(aw[r][1], __funnelshift_rc(lo.y, lo.z, sh), t);
                    t = __dp4a(aw[r][2], __funnelshift_rc(lo.z, lo.w, sh), t);
                    acc[dy * 3 + dx] = __dp4a(aw[r][3], __funnelshift_rc(lo.w, w4, sh), t);
                }
            }
        }
    }
    // butterfly over the 8 lanes of the child: 16 -> 2 values per lane (14 shuffles);
    // lane bits (b2 b1 b0) end with candidates 8*b2 + 4*b1 + 2*b0 and the next one
    uint32_t r0v, r1v;
    {
        uint32_t w8[8], w4[4];
        const bool h4 = l & 4, h2 = l & 2, h1 = l & 1;
#pragma unroll
        for (int q = 0; q < 8; ++q) w8[q] = (h4 ? acc[q + 8] : acc[q]) + __shfl_xor_sync(0xffffffffu, h4 ? acc[q] : acc[q + 8], 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) w4[q] = (h2 ? w8[q + 4] : w8[q]) + __shfl_xor_sync(0xffffffffu, h2 ? w8[q] : w8[q + 4], 2);
        r0v = (h1 ? w4[2] : w4[0]) + __shfl_xor_sync(0xffffffffu, h1 ? w4[0] : w4[2], 1);
        r1v = (h1 ? w4[3] : w4[1]) + __shfl_xor_sync(0xffffffffu, h1 ? w4[1] : w4[3], 1);
    }
    const int s_base = ((l >> 2) & 1) * 8 + ((l >> 1) & 1) * 4 + (l & 1) * 2;
    float val0 = 0.0f, val1 = 0.0f;
    {
        const int sA = s_base, sB = s_base + 1;
        const int qyA = d0 + sA / 3 - 1, qxA = d1 + sA % 3 - 1, qyB = d0 + sB / 3 - 1, qxB = d1 + sB % 3 - 1;
        if (sA < 9 && qyA >= 0 && qyA < T0 && qxA >= 0 && qxA < T1) val0 = value_at((int)r0v, qyA, qxA);
        if (sB < 9 && qyB >= 0 && qyB < T0 && qxB >= 0 && qxB < T1) val1 = value_at((int)r1v, qyB, qxB);
    }
    float best = 0.f, centre = 0.f;
    int bi = 0;
    bool best_nan = false;
#pragma unroll
    for (int s = 0; s < 9; ++s) {
        const int src = ((s >> 3) & 1) * 4 + ((s >> 2) & 1) * 2 + ((s >> 1) & 1);
        const float v = __shfl_sync(0xffffffffu, (s & 1) ? val1 : val0, gbase + src);
        if (s == 4) centre = v;
        if (s == 0) { best = v; best_nan = (v != v); }
        else if (!best_nan && (v > best || v != v)) { best = v; bi = s; best_nan = (v != v); }
    }
    // level-0 values (co_map ** 1.4) of the winner and of the centre
    best = dm_rectify(best); centre = dm_rectify(centre);
    if (best < DM_NEAR_ZERO_F) { bi = 4; best = centre; }
    const int c0 = d0 + bi / 3 - 1, c1 = d1 + bi % 3 - 1;
    const float score = best + centre;

    // ---- misc/Matching.py:165-209 parabola fit (index -1 wraps, upper edge skipped)
    double mrow = (double)c0, mcol = (double)c1;
    if (a.sub_pix) {                                     // warp-uniform
        const bool in = c0 >= 0 && c0 < T0 && c1 >= 0 && c1 < T1;
        // neighbours: 0 = (c0+1,c1)  1 = (c0-1 | wrap, c1)  2 = (c0,c1+1)  3 = (c0, c1-1 | wrap)
        const int ny[4] = {c0 + 1, (c0 == 0 ? T0 - 1 : c0 - 1), c0, c0};
        const int nx[4] = {c1, c1, c1 + 1, (c1 == 0 ? T1 - 1 : c1 - 1)};
        const bool nok[4] = {in && c0 + 1 < T0, in && c0 + 1 < T0, in && c1 + 1 < T1, in && c1 + 1 < T1};
        uint32_t nacc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            nacc[q] = 0;
            if (!nok[q]) continue;
            // window start relative to the child's candidate origin (d0 - 2, d1 - 2)
            const int ry = ny[q] - (d0 - 2), rx = nx[q] - (d1 - 2);
            const bool inreg = ry >= 0 && ry <= 4 && rx >= 0 && rx <= 4;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int ky = l + 8 * r;
                if (ky >= WS) continue;
                if (inreg) {
                    nacc[q] += row_dot(aw[r], ky + ci + ry, rx);
                } else {                                 // wrapped index: far away, read global memory
                    const uint8_t* b = a.img2 + (size_t)(oy + ny[q] + ky) * a.pitch + ox + nx[q];
                    uint32_t bw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (u < WS) bw[u >> 2] |= (uint32_t)__ldg(b + u) << (8 * (u & 3));
                    nacc[q] = __dp4a(aw[r][3], bw[3], __dp4a(aw[r][2], bw[2], __dp4a(aw[r][1], bw[1], __dp4a(aw[r][0], bw[0], nacc[q]))));
                }
            }
        }
        // 4 partial sums over 8 lanes -> lane holds neighbour 2*b2 + b1 (4 shuffles)
        const bool h4 = l & 4, h2 = l & 2;
        uint32_t u0 = (h4 ? nacc[2] : nacc[0]) + __shfl_xor_sync(0xffffffffu, h4 ? nacc[0] : nacc[2], 4);
        uint32_t u1 = (h4 ? nacc[3] : nacc[1]) + __shfl_xor_sync(0xffffffffu, h4 ? nacc[1] : nacc[3], 4);
        uint32_t tot = (h2 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, h2 ? u0 : u1, 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        const int my_n = ((l >> 2) & 1) * 2 + ((l >> 1) & 1);
        const int sy = my_n == 0 ? ny[0] : (my_n == 1 ? ny[1] : c0);
        const int sx = my_n == 2 ? nx[2] : (my_n == 3 ? nx[3] : c1);
        const bool sok = my_n < 2 ? nok[0] : nok[2];
        float nval = 0.0f;
        if (sok) nval = dm_rectify(value_at((int)tot, sy, sx));
        const float r0 = best;                      // level-0 value at the match itself
        const float v0 = __shfl_sync(0xffffffffu, nval, gbase + 0), v1 = __shfl_sync(0xffffffffu, nval, gbase + 2);
        const float v2 = __shfl_sync(0xffffffffu, nval, gbase + 4), v3 = __shfl_sync(0xffffffffu, nval, gbase + 6);
        if (nok[0] && r0 > v0 && r0 > v1) mrow += dm_parabola_shift(r0, v0, v1);
        if (nok[2] && r0 > v2 && r0 > v3) mcol += dm_parabola_shift(r0, v2, v3);
    }
    if (l != 0) return;

    // ---- planes + paste (misc/Calc_difference.py:36-48, misc/image_cut_solver.py:165-175)
    const int gi = ti.x, gj = ti.y, sc = ti.z;
    const int Y = a.s0 * gi + i, X = a.s1 * gj + j;
    // owner = min(Y / s0, len0 - 1) with Y / s0 = gi + i / s0 (same for the columns)
    if (min(gi + (int)dm_fd_div((uint32_t)i, a.fd_s0), a.len0 - 1) != gi || min(gj + (int)dm_fd_div((uint32_t)j, a.fd_s1), a.len1 - 1) != gj) return;   // a later tile owns this pixel
    const double e0 = __dsub_rn((double)i, mrow), e1 = __dsub_rn((double)j, mcol);
    const size_t plane = (size_t)a.out_h * a.out_w, pix = (size_t)Y * a.out_w + X;
    for (int m = 0; m < a.n_modes; ++m) {
        const double v = a.modes[m] == DM_MODE_ELEVATION ? e1
                       : a.modes[m] == DM_MODE_ELEVATION2 ? e0
                       : __dsqrt_rn(__fma_rn(e1, e1, __dmul_rn(e0, e0)));
        a.d_map[((size_t)sc * a.n_modes + m) * plane + pix] = v;
    }
    a.out_map[(size_t)sc * plane + pix] = (double)score;
}

// ---------------------------------------------------------------------------------------
// Final level, one THREAD per patch (round 2).  The warp-per-quad kernel above spends ~1300 warp
// instructions per quad (7300 thread instructions per patch at ws 15, the same at ws 5) on staging,
// shuffles and reductions around 13 window correlations that are 13 x ws x ceil(ws/4) DP4As.  Here a
// thread keeps its own patch in registers (ws rows of packed bytes, fetched as aligned 32-bit words +
// one funnel shift each) and walks every candidate window of image 2 the same way: no shared memory,
// no shuffles, no reductions; neighbouring threads are neighbouring patches, so their loads fall into
// the same cache lines and their plane stores are contiguous.  Same arithmetic, same order of the
// comparisons: bit-identical to the kernel above (tests/test_gpu_parity.py).
// misc/Matching.py:58-78,98-139,165-209 + misc/Calc_difference.py:36-48 + the paste of
// misc/image_cut_solver.py:165-175.
// ---------------------------------------------------------------------------------------
template <int WS>
__global__ void __launch_bounds__(128)
dm_final_patch_kernel(const FinalArgs a, long long n_patches) {
    constexpr int K = WS * WS;
    constexpr int NW = (WS + 3) / 4;                       // packed words per window row
    constexpr uint32_t LASTMASK = (WS % 4) ? ((1u << (8 * (WS % 4))) - 1u) : 0xffffffffu;
    dm_pdl_wait();                  // before the early exits: the grid must not complete ahead of the kernel in front
    dm_pdl_launch_dependents();
    const long long tl = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tl >= n_patches) return;
    const uint32_t t = (uint32_t)tl;                       // a launch starts at a tile boundary and never holds 2^32 patches
    const int T0 = a.t0, T1 = a.t1, P = T0 * T1;
    const uint32_t nrel = dm_fd_div(t, a.fd_p);
    const int n = a.tile0 + (int)nrel;                     // tile inside the chunk
    const int p = (int)(t - nrel * (uint32_t)P);
    const int i = (int)dm_fd_div((uint32_t)p, a.fd_t1), j = p - i * T1;
    const int4 ti = __ldg(reinterpret_cast<const int4*>(a.tinfo) + n);      // (gi, gj, scene, -)
    const int gi = ti.x, gj = ti.y, sc = ti.z;
    // misc/image_cut_solver.py:165-175: a pixel belongs to the covering tile with the largest index
    if (min(gi + (int)dm_fd_div((uint32_t)i, a.fd_s0), a.len0 - 1) != gi || min(gj + (int)dm_fd_div((uint32_t)j, a.fd_s1), a.len1 - 1) != gj) return;
    const int oy = a.origin[2 * n], ox = a.origin[2 * n + 1];
    const bool normed = a.normed != 0;
    const int hA = T0 >> 1, hB = T1 >> 1, PQ = hA * hB;
    // misc/Matching.py:116-124: p_dot = 2 * parent match + o
    const int pq = (i >> 1) * hB + (j >> 1);
    const size_t pb = (size_t)n * 2 * PQ;
    const int d0 = 2 * a.parent[pb + pq] + (i & 1), d1 = 2 * a.parent[pb + PQ + pq] + (j & 1);

    const dm_stat s1 = a.stat1[(size_t)n * P + p];
    const bool flat1 = (s1.y == 0.0f);
    const float4 pmn = reinterpret_cast<const float4*>(a.rowmin)[(size_t)n * P + p], pmx = reinterpret_cast<const float4*>(a.rowmax)[(size_t)n * P + p];
    const float mn = dm_min_nan(dm_min_nan(pmn.x, pmn.y), dm_min_nan(pmn.z, pmn.w));
    const float mx = dm_max_nan(dm_max_nan(pmx.x, pmx.y), dm_max_nan(pmx.z, pmx.w)), rinv = dm_range_inv(mn, mx);
    const dm_stat* st2 = a.stat2 + (size_t)n * P;
    const int m1 = (int)s1.w, S1 = (int)s1.x;

    // WS bytes starting at `src` as NW packed words: aligned 32-bit loads (only words that hold a needed
    // byte) + funnel shifts, bytes beyond the window masked off
    auto load_row = [&](const uint8_t* src, uint32_t (&out)[NW]) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
        const int off = (int)(addr & 3);
        uint32_t w[NW + 1];
#pragma unroll
        for (int k = 0; k <= NW; ++k) w[k] = (4 * k < off + WS) ? __ldg(ap + k) : 0u;
#pragma unroll
        for (int k = 0; k < NW; ++k) out[k] = __funnelshift_r(w[k], w[k + 1], 8 * off);
        out[NW - 1] &= LASTMASK;
    };
    uint32_t aw[WS][NW];                                   // this thread's patch
    {
        const uint8_t* b1 = a.img1 + (size_t)(oy + i) * a.pitch + ox + j;
#pragma unroll
        for (int ky = 0; ky < WS; ++ky) load_row(b1 + (size_t)ky * a.pitch, aw[ky]);
    }
    // min-maxed (not yet rectified) co_map value of position (qy,qx): exact integer sum a*b over the window,
    // then the same formula as everywhere else (see dm_final_quad_kernel::value_at)
    auto value_at = [&](int qy, int qx) -> float {
        const uint8_t* b2 = a.img2 + (size_t)(oy + qy) * a.pitch + ox + qx;
        uint32_t acc = 0;
#pragma unroll
        for (int ky = 0; ky < WS; ++ky) {
            uint32_t bw[NW];
            load_row(b2 + (size_t)ky * a.pitch, bw);
#pragma unroll
            for (int k = 0; k < NW; ++k) acc = __dp4a(aw[ky][k], bw[k], acc);
        }
        const dm_stat sq = st2[qy * T1 + qx];
        const int m2 = (int)sq.w, S2 = (int)sq.x;
        const int dot = (int)acc - m2 * S1 - m1 * S2 - K * m1 * m2;
        const float z = dm_zncc_partial((float)dot, s1.x, sq.z, normed ? sq.y : 1.0f);
        const float r = dm_normalize(dm_zncc_finish(z, s1.y, flat1, normed), mn, mx, rinv);
        return r < 0.0f ? 0.0f : r;                        // see dm_final_quad_kernel
    };

    // ---- misc/Matching.py:58-78 on the recomputed 3x3 window (zero padding outside)
    float best = 0.f, centre = 0.f;
    int bi = 0;
    bool best_nan = false;
#pragma unroll 1
    for (int s = 0; s < 9; ++s) {
        const int qy = d0 + s / 3 - 1, qx = d1 + s % 3 - 1;
        float v = 0.0f;
        if (qy >= 0 && qy < T0 && qx >= 0 && qx < T1) v = value_at(qy, qx);
        if (s == 4) centre = v;
        if (s == 0) { best = v; best_nan = (v != v); }
        else if (!best_nan && (v > best || v != v)) { best = v; bi = s; best_nan = (v != v); }
    }
    best = dm_rectify(best); centre = dm_rectify(centre);
    if (best < DM_NEAR_ZERO_F) { bi = 4; best = centre; }
    const int c0 = d0 + bi / 3 - 1, c1 = d1 + bi % 3 - 1;
    const float score = best + centre;

    // ---- misc/Matching.py:165-209 parabola fit (index -1 wraps, upper edge skipped)
    double mrow = (double)c0, mcol = (double)c1;
    if (a.sub_pix) {
        const bool in = c0 >= 0 && c0 < T0 && c1 >= 0 && c1 < T1;
        const float r0 = best;
        if (in && c0 + 1 < T0) {
            const float v0 = dm_rectify(value_at(c0 + 1, c1)), v1 = dm_rectify(value_at(c0 == 0 ? T0 - 1 : c0 - 1, c1));
            if (r0 > v0 && r0 > v1) mrow += dm_parabola_shift(r0, v0, v1);
        }
        if (in && c1 + 1 < T1) {
            const float v2 = dm_rectify(value_at(c0, c1 + 1)), v3 = dm_rectify(value_at(c0, c1 == 0 ? T1 - 1 : c1 - 1));
            if (r0 > v2 && r0 > v3) mcol += dm_parabola_shift(r0, v2, v3);
        }
    }

    // ---- planes + paste (misc/Calc_difference.py:36-48, misc/image_cut_solver.py:165-175)
    const int Y = a.s0 * gi + i, X = a.s1 * gj + j;
    const double e0 = __dsub_rn((double)i, mrow), e1 = __dsub_rn((double)j, mcol);
    const size_t plane = (size_t)a.out_h * a.out_w, pix = (size_t)Y * a.out_w + X;
    for (int m = 0; m < a.n_modes; ++m) {
        const double v = a.modes[m] == DM_MODE_ELEVATION ? e1
                       : a.modes[m] == DM_MODE_ELEVATION2 ? e0
                       : __dsqrt_rn(__fma_rn(e1, e1, __dmul_rn(e0, e0)));
        a.d_map[((size_t)sc * a.n_modes + m) * plane + pix] = v;
    }
    a.out_map[(size_t)sc * plane + pix] = (double)score;
}

template <int WS>
static void launch_final_patch(const FinalArgs& fa, long long n_patches, cudaStream_t st) {
    dm_launch_dep(DM_PDL_FINAL, dm_final_patch_kernel<WS>, dim3((unsigned)dm_div_up(n_patches, 128)), dim3(128), 0, st, fa, n_patches);
}

template <int WS>
static void launch_final_quad(const FinalArgs& fa, long long n_quads, cudaStream_t st) {
    dm_final_quad_kernel<WS><<<dm_div_up(n_quads, FQ_WARPS), 32 * FQ_WARPS, 0, st>>>(fa, n_quads);
}

}  // namespace

bool dm_fused_supported(int t0, int t1, int kpad) {
    return dm_correlation_umma_pool_supported(t0, t1, kpad) && t0 >= 4 && t1 >= 4;
}
bool dm_fused_supported_ws(int ws) { return ws >= 3 && ws <= 15 && (ws & 1); }

size_t dm_fused_workspace(char* base, int n_tiles, int t0, int t1, int kpad, int levels, void* buffers_out) {
    FusedBuffers fb;
    size_t n = carve(base, n_tiles, t0, t1, kpad, levels, fb);
    if (buffers_out) *(FusedBuffers*)buffers_out = fb;
    return n;
}

int dm_fused_solve_chunk(dm_ctx* ctx, const dm_fused_args* a, int ck) {
    cudaStream_t st = ctx->stream;
    FusedBuffers fb;
    const int nt = a->n_tiles, t0 = a->t0, t1 = a->t1, P = t0 * t1, L = a->levels;
    carve(ctx->ws, nt, t0, t1, a->kpad, L, fb);
    int rc;
    {
        StageTimer tm(ctx, DM_STAGE_DESCRIPTORS);
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        dm_tile_origin_kernel2<<<dm_div_up(nt, 128), 128, 0, st>>>(fb.origin, fb.tinfo, nt, a->first_tile, a->len0, a->len1, a->s0, a->s1, a->scene_h);
        DM_LAUNCH_CHECK();
        int n_desc = 0;
        if ((rc = dm_descriptors_both(a->img1, a->img2, a->scene_h * a->n_scenes, a->scene_w, a->scene_w, fb.origin, nt, t0, t1, a->ws,
                                      fb.desc1, fb.stat1, fb.desc2, fb.stat2, st, &n_desc)) != DM_OK) return rc;
        ctx->launches[DM_STAGE_DESCRIPTORS] += 1 + n_desc;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    {
        StageTimer tm(ctx, DM_STAGE_CORRELATION);
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        if ((rc = dm_correlation_umma_pool(fb.desc1, fb.stat1, fb.desc2, fb.stat2, nt, t0, t1, a->kpad, dm_desc_kreal(a->ws), a->method,
                                           0, fb.pooled, fb.rowmin, fb.rowmax, st)) != DM_OK) return rc;
        ctx->launches[DM_STAGE_CORRELATION] += 1;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    {
        StageTimer tm(ctx, DM_STAGE_NORMALIZE);        // min-max + rectify + first child average
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        const long long parents = (long long)nt * (P / 4);
        const int q4 = P / 16;
        // 64 threads making four or more trips each, not 256 threads and one trip: 32 small CTAs stay
        // resident per SM and a CTA's loads overlap its own arithmetic (C2: 0.838 -> 0.676 ms, 7.0 TB/s)
        const int threads = q4 >= 256 ? 64 : (q4 < 32 ? 32 : q4);
        static const bool first_cta = getenv("DM_FIRST_CTA") != nullptr;       // measurement aid: one CTA per parent for small maps too
        if (q4 <= 64 && !first_cta) {
            // small maps: a persistent grid of warps striding over the parents (dm_aggregate_first_small_kernel)
            int dev = 0, sms = 148, per_sm = 4;
            auto sk = q4 > 32 ? dm_aggregate_first_small_kernel<2> : dm_aggregate_first_small_kernel<1>;
            DM_CUDA_CHECK(cudaGetDevice(&dev));
            DM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            DM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sk, 256, 0));
            long long blocks = (long long)sms * (per_sm > 0 ? per_sm : 1);
            if (blocks > (parents + 7) / 8) blocks = (parents + 7) / 8;
            DM_CUDA_CHECK(dm_launch_dep(DM_PDL_FIRST, sk, dim3((unsigned)blocks), dim3(256), 0, st,
                                        (const float*)fb.pooled, (const float*)fb.rowmin, (const float*)fb.rowmax, t0, t1, (unsigned)parents,
                                        dm_make_fastdiv((uint32_t)(t1 >> 1)), dm_make_fastdiv((uint32_t)(t0 >> 1)), fb.level[1]));
        } else {
            DM_CUDA_CHECK(dm_launch_dep(DM_PDL_FIRST, dm_aggregate_first_kernel, dim3((unsigned)parents), dim3(threads), 0, st,
                                        (const float*)fb.pooled, (const float*)fb.rowmin, (const float*)fb.rowmax, t0, t1, fb.level[1]));
        }
        DM_LAUNCH_CHECK();
        ctx->launches[DM_STAGE_NORMALIZE] += 1;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    // ---- upper pyramid + top-down pass: up to 2 (levels - 1) small dependent launches that touch
    // nothing but the workspace.  They are captured once per (workspace, shape, filter) into a CUDA
    // graph (on a private stream: the caller's may be the legacy default stream, which cannot
    // capture) and replayed with one launch; with stage timing on they are launched one by one.
    int cur = 0;
    // without the displacement filter the levels above ks and the whole top-down pass run as ONE
    // kernel (one CTA per tile, dm_upper_tail); the generic kernels then only build levels 2 .. ks
    static const bool no_tail = getenv("DM_NO_TAIL") != nullptr;
    const bool tail = !no_tail && a->filter_num <= 0 && dm_upper_tail_supported(t0, t1, L);
    const int agg_end = tail ? dm_upper_tail_first_level(t0, t1, L) : L - 1;          // last level built by dm_aggregate
    auto run_agg = [&](cudaStream_t s, int* n_agg) -> int {
        *n_agg = 0;
        for (int k = 1; k + 1 <= agg_end; ++k) {
            int r = dm_aggregate(fb.level[k], nt, t0 >> k, t1 >> k, t0 >> k, t1 >> k, 1, fb.level[k + 1], s);
            if (r != DM_OK) return r;
            ++*n_agg;
        }
        return DM_OK;
    };
    auto run_bt = [&](cudaStream_t s, int* n_bt) -> int {
        int r;
        cur = 0; *n_bt = 0;
        if (tail) {
            if ((r = dm_upper_tail(fb.level, nt, t0, t1, L, fb.match[0], s)) != DM_OK) return r;
            *n_bt = 1;
            return DM_OK;
        }
        int filters_left = a->filter_num;
        auto maybe_filter = [&](int k) -> int {       // misc/Matching.py:91-93,136-138
            if (filters_left <= 0) return DM_OK;
            --filters_left;
            const int h = t0 >> k, w = t1 >> k;
            if (!(h >= a->filter_win && w >= a->filter_win)) return DM_OK;
            int rr = dm_match_filter(fb.match[cur], nt, h, w, a->filter_win, a->filter_mode, fb.match[cur ^ 1], s);
            if (rr != DM_OK) return rr;
            cur ^= 1;
            ++*n_bt;
            return DM_OK;
        };
        // the scores of the levels above 0 are never read on this path (the final kernel
        // recomputes the level-0 score), so the score buffers simply follow the match index
        if ((r = dm_backtrack_top(fb.level[L - 1], 0, nt, t0 >> (L - 1), t1 >> (L - 1), fb.match[cur], fb.score[cur], s)) != DM_OK) return r;
        ++*n_bt;
        if ((r = maybe_filter(L - 1)) != DM_OK) return r;
        for (int k = L - 2; k >= 1; --k) {
            if ((r = dm_backtrack_level(fb.level[k], 0, nt, t0 >> k, t1 >> k, t0 >> k, t1 >> k, fb.match[cur], fb.match[cur ^ 1], fb.score[cur ^ 1], s)) != DM_OK) return r;
            cur ^= 1;
            ++*n_bt;
            if ((r = maybe_filter(k)) != DM_OK) return r;
        }
        return DM_OK;
    };
    if (ctx->timing) {
        int n = 0;
        {
            StageTimer tm(ctx, DM_STAGE_AGGREGATE);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            if ((rc = run_agg(st, &n)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_AGGREGATE] += n;
            if ((rc = tm.end()) != DM_OK) return rc;
        }
        {
            StageTimer tm(ctx, DM_STAGE_BACKTRACK);
            if ((rc = tm.begin(ck)) != DM_OK) return rc;
            if ((rc = run_bt(st, &n)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_BACKTRACK] += n;
            if ((rc = tm.end()) != DM_OK) return rc;
        }
    } else {
        dm_ctx::UpperGraph* ug = nullptr;
        for (auto& g : ctx->upper_graphs)
            if (g.ws == (const void*)ctx->ws && g.nt == nt && g.t0 == t0 && g.t1 == t1 && g.levels == L && g.kpad == a->kpad &&
                g.filter_num == a->filter_num && g.filter_win == a->filter_win && g.filter_mode == a->filter_mode) { ug = &g; break; }
        if (!ug) {
            if (ctx->upper_graphs.size() >= 8) {          // a scene uses at most two chunk sizes per geometry
                for (auto& g : ctx->upper_graphs) cudaGraphExecDestroy(g.exec);
                ctx->upper_graphs.clear();
            }
            if (!ctx->capture_stream) DM_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
            dm_ctx::UpperGraph g;
            g.ws = ctx->ws; g.nt = nt; g.t0 = t0; g.t1 = t1; g.levels = L; g.kpad = a->kpad;
            g.filter_num = a->filter_num; g.filter_win = a->filter_win; g.filter_mode = a->filter_mode;
            // the launches carry the programmatic-dependent-launch attribute (programmatic edges between the kernel
            // nodes); should a driver refuse those in a capture, the sequence is captured once more without it
            cudaError_t ce = cudaSuccess;
            for (int attempt = 0; attempt < 2; ++attempt) {
                DM_CUDA_CHECK(cudaStreamBeginCapture(ctx->capture_stream, cudaStreamCaptureModeThreadLocal));
                rc = run_agg(ctx->capture_stream, &g.n_agg);
                if (rc == DM_OK) rc = run_bt(ctx->capture_stream, &g.n_bt);
                cudaGraph_t graph = nullptr;
                ce = cudaStreamEndCapture(ctx->capture_stream, &graph);
                if (rc == DM_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&g.exec, graph, 0);
                if (graph) cudaGraphDestroy(graph);
                if (rc == DM_OK && ce == cudaSuccess) break;
                if (attempt == 0 && dm_pdl_enabled(DM_PDL_UPPER)) { cudaGetLastError(); dm_pdl_suppress(true); rc = DM_OK; continue; }
                break;
            }
            dm_pdl_suppress(false);
            if (rc != DM_OK) return rc;
            DM_CUDA_CHECK(ce);
            g.final_cur = cur;
            ctx->upper_graphs.push_back(g);
            ug = &ctx->upper_graphs.back();
        }
        DM_CUDA_CHECK(cudaGraphLaunch(ug->exec, st));
        cur = ug->final_cur;
        ctx->launches[DM_STAGE_AGGREGATE] += ug->n_agg;
        ctx->launches[DM_STAGE_BACKTRACK] += ug->n_bt;
    }
    {
        StageTimer tm(ctx, DM_STAGE_PLANES);           // level-0 backtracking + sub-pixel + planes
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        FinalArgs fa;
        fa.img1 = a->img1; fa.img2 = a->img2; fa.pitch = a->scene_w;
        fa.origin = fb.origin; fa.tinfo = fb.tinfo;
        fa.fd_pq = dm_make_fastdiv((uint32_t)(P / 4)); fa.fd_hb = dm_make_fastdiv((uint32_t)(t1 >> 1));
        fa.fd_s0 = dm_make_fastdiv((uint32_t)a->s0); fa.fd_s1 = dm_make_fastdiv((uint32_t)a->s1);
        fa.fd_p = dm_make_fastdiv((uint32_t)P); fa.fd_t1 = dm_make_fastdiv((uint32_t)t1);
        fa.stat1 = (const dm_stat*)fb.stat1; fa.stat2 = (const dm_stat*)fb.stat2;
        fa.rowmin = fb.rowmin; fa.rowmax = fb.rowmax; fa.parent = fb.match[cur];
        fa.t0 = t0; fa.t1 = t1; fa.ws = a->ws; fa.normed = a->method == DM_TM_CCOEFF_NORMED; fa.sub_pix = a->sub_pix;
        fa.n_modes = a->n_modes; for (int m = 0; m < 4; ++m) fa.modes[m] = a->modes[m];
        fa.s0 = a->s0; fa.s1 = a->s1; fa.len0 = a->len0; fa.len1 = a->len1; fa.out_h = a->out_h; fa.out_w = a->out_w;
        fa.first_tile = a->first_tile; fa.d_map = a->d_map; fa.out_map = a->out_map;
        const long long n_patches = (long long)nt * P;
        fa.scene_h = a->scene_h * a->n_scenes;       // bound of the stacked image
        const long long n_quads = n_patches / 4;
        // streaming active: the stage runs in (up to) four bands of tiles, and what each band's tiles own starts
        // its copy while the next band runs
        int bands = 1;
        if (ctx->rb.active) { bands = nt / (2 * a->len1); if (bands > 4) bands = 4; if (bands < 1) bands = 1; }   // only the last band's copy is exposed
        int t_begin = 0;
        for (int b = 0; b < bands; ++b) {
            int t_end = (b == bands - 1) ? nt : (int)((long long)nt * (b + 1) / bands);
            if (t_end <= t_begin) continue;
            fa.tile0 = t_begin;
            const long long nq = (long long)(t_end - t_begin) * (P / 4);
            static const bool per_quad = getenv("DM_FINAL_QUAD") != nullptr;      // measurement aid: the warp-per-quad kernel
            if (!per_quad) switch (a->ws) {
                case 3: launch_final_patch<3>(fa, nq * 4, st); break;
                case 5: launch_final_patch<5>(fa, nq * 4, st); break;
                case 7: launch_final_patch<7>(fa, nq * 4, st); break;
                case 9: launch_final_patch<9>(fa, nq * 4, st); break;
                case 11: launch_final_patch<11>(fa, nq * 4, st); break;
                case 13: launch_final_patch<13>(fa, nq * 4, st); break;
                case 15: launch_final_patch<15>(fa, nq * 4, st); break;
                default: DM_REQUIRE(false, DM_ERR_UNSUPPORTED, "fused path supports odd window sizes 3..15 (got %d)", a->ws);
            }
            else switch (a->ws) {
                case 3: launch_final_quad<3>(fa, nq, st); break;
                case 5: launch_final_quad<5>(fa, nq, st); break;
                case 7: launch_final_quad<7>(fa, nq, st); break;
                case 9: launch_final_quad<9>(fa, nq, st); break;
                case 11: launch_final_quad<11>(fa, nq, st); break;
                case 13: launch_final_quad<13>(fa, nq, st); break;
                case 15: launch_final_quad<15>(fa, nq, st); break;
                default: DM_REQUIRE(false, DM_ERR_UNSUPPORTED, "fused path supports odd window sizes 3..15 (got %d)", a->ws);
            }
            DM_LAUNCH_CHECK();
            ctx->launches[DM_STAGE_PLANES] += 1;
            if (ctx->rb.active && (rc = dm_readback_rows(ctx, (long long)a->first_tile + t_end)) != DM_OK) return rc;
            t_begin = t_end;
        }
        (void)n_quads;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    return DM_OK;
}
