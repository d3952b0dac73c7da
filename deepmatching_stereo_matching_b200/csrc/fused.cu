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
#include "dm_common.cuh"
#include "dm_internal.h"

namespace {

struct FusedBuffers {
    int32_t* origin; void* desc1; void* desc2; float* stat1; float* stat2;
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
    fb.desc1 = c.take<__nv_bfloat16>((size_t)nt * P * kpad);
    fb.desc2 = c.take<__nv_bfloat16>((size_t)nt * P * kpad);
    fb.stat1 = c.take<float>((size_t)nt * P * DM_STAT_FLOATS);
    fb.stat2 = c.take<float>((size_t)nt * P * DM_STAT_FLOATS);
    fb.pooled = c.take<float>((size_t)nt * P * (P / 4));
    fb.rowmin = c.take<float>((size_t)nt * P);
    fb.rowmax = c.take<float>((size_t)nt * P);
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

__global__ void dm_tile_origin_kernel2(int32_t* origin, int n, int first_tile, int len1, int s0, int s1) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int g = first_tile + t;
    origin[2 * t] = s0 * (g / len1);
    origin[2 * t + 1] = s1 * (g % len1);
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
        mn[ch] = __ldg(rowmin + p); mx[ch] = __ldg(rowmax + p); rinv[ch] = dm_range_inv(mn[ch], mx[ch]);
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

// ---------------------------------------------------------------------------------------
// Final level.  One warp per patch p = (i,j) of a tile.
// ---------------------------------------------------------------------------------------
struct FinalArgs {
    const uint8_t* img1; const uint8_t* img2; int pitch;
    const int32_t* origin; const dm_stat* stat1; const dm_stat* stat2;
    const float* rowmin; const float* rowmax;
    const int32_t* parent;      // level-1 matches [n][2][t0/2][t1/2]
    int t0, t1, ws, normed, sub_pix, scene_h;
    int n_modes, modes[4];
    int s0, s1, len0, len1, out_h, out_w, first_tile;
    double* d_map; double* out_map;
};

// The 3x3 candidate windows around p_dot and the parabola neighbours of the match all lie
// in the (ws+4)^2 pixel region of image 2 that starts two pixels up/left of window p_dot.
// The warp stages that region in shared memory once (word-aligned rows).  Lane (ky, hf)
// owns bytes [8*hf, 8*hf+8) of patch row ky as two packed words and sweeps the shifted
// windows with funnel shifts + DP4A (4 byte-MACs per instruction).  All sums are exact:
//   sum (a-m1)(b-m2) = sum a*b - m2*S1' - m1*S2' - K*m1*m2     (S' = residual sums of the stats)
constexpr int RSB = 24;                 // region row stride in bytes (ws + 4 <= 19, word aligned)
constexpr int RBYTES = 19 * RSB + 8;

template <int WS>
__global__ void __launch_bounds__(256)
dm_final_level_kernel(const FinalArgs a, long long n_patches) {
    constexpr int K = WS * WS;
    constexpr int RW = WS + 4;
    __shared__ __align__(16) uint8_t region_all[8][RBYTES];
    const int lane = threadIdx.x & 31;
    uint8_t* region = region_all[threadIdx.x >> 5];
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_patches) return;
    const int T0 = a.t0, T1 = a.t1, P = T0 * T1;
    const int n = (int)(w / P);                          // tile inside the chunk
    const int p = (int)(w - (long long)n * P);
    const int i = p / T1, j = p - i * T1;
    const int oy = a.origin[2 * n], ox = a.origin[2 * n + 1];
    const bool normed = a.normed != 0;

    // misc/Matching.py:116-124: p_dot = 2 * parent match + o
    const int hA = T0 >> 1, hB = T1 >> 1;
    const size_t pb = (size_t)n * 2 * hA * hB, pi = (size_t)(i >> 1) * hB + (j >> 1);
    const int d0 = 2 * a.parent[pb + pi] + (i & 1);
    const int d1 = 2 * a.parent[pb + (size_t)hA * hB + pi] + (j & 1);

    const dm_stat s1 = a.stat1[(size_t)n * P + p];
    const bool flat1 = (s1.y == 0.0f);
    const float mn = a.rowmin[(size_t)n * P + p], mx = a.rowmax[(size_t)n * P + p], rinv = dm_range_inv(mn, mx);
    const dm_stat* st2 = a.stat2 + (size_t)n * P;
    const int m1 = (int)s1.w, S1 = (int)s1.x;

    // stage the region: rows oy+d0-2 .. +RW, cols ox+d1-2 .. +RW (zeros outside the scene).
    // Fully unrolled so that all RW row loads of a lane are in flight together; the common
    // case (region entirely inside the scene) runs without per-row guards.
    {
        const int gy0 = oy + d0 - 2, gx0 = ox + d1 - 2;
        const int gx = gx0 + lane;
        uint8_t vals[RW];
        if (gy0 >= 0 && gy0 + RW <= a.scene_h && gx0 >= 0 && gx0 + RW <= a.pitch) {     // warp-uniform
            const uint8_t* src = a.img2 + (size_t)gy0 * a.pitch + gx0 + (lane < RW ? lane : 0);
#pragma unroll
            for (int ry = 0; ry < RW; ++ry) { vals[ry] = __ldg(src); src += a.pitch; }
        } else {
            const bool colok = lane < RW && gx >= 0 && gx < a.pitch;
#pragma unroll
            for (int ry = 0; ry < RW; ++ry) {
                const int gy = gy0 + ry;
                vals[ry] = (colok && gy >= 0 && gy < a.scene_h) ? __ldg(a.img2 + (size_t)gy * a.pitch + gx) : (uint8_t)0;
            }
        }
        if (lane < RSB) {
            uint8_t* dst = region + lane;
#pragma unroll
            for (int ry = 0; ry < RW; ++ry) dst[ry * RSB] = (lane < RW) ? vals[ry] : (uint8_t)0;
        }
    }
    // this lane's 8 patch bytes (row ky, columns 8*hf .. 8*hf+7; zero beyond the window)
    const int ky = lane >> 1, hf = lane & 1;
    const bool rowlive = ky < WS;
    uint32_t a0 = 0, a1 = 0;
    {
        const uint8_t* a_row = a.img1 + (size_t)(oy + i + (rowlive ? ky : 0)) * a.pitch + ox + j + hf * 8;
        const int nvalid = rowlive ? (hf ? WS - 8 : (WS < 8 ? WS : 8)) : 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t v = (u < nvalid) ? (uint32_t)__ldg(a_row + u) : 0u;
            if (u < 4) a0 |= v << (8 * u); else a1 |= v << (8 * (u - 4));
        }
    }
    __syncwarp();
    const uint32_t* rwords = reinterpret_cast<const uint32_t*>(region) + hf * 2;    // lane's first word in a region row

    // level-0 value of position (qy,qx) from sum a*b (each lane evaluates ONE candidate)
    auto value_at = [&](int sum_ab, int qy, int qx) -> float {
        const dm_stat sq = st2[qy * T1 + qx];
        const int m2 = (int)sq.w, S2 = (int)sq.x;
        const int dot = sum_ab - m2 * S1 - m1 * S2 - K * m1 * m2;
        const float z = dm_zncc_partial((float)dot, s1.x, sq.z, normed ? sq.y : 1.0f);
        return dm_rectify(dm_normalize(dm_zncc_finish(z, s1.y, flat1, normed), mn, mx, rinv));
    };

    // ---- misc/Matching.py:58-78 on the recomputed 3x3 window (zero padding outside)
    uint32_t acc[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) acc[s] = 0;
    if (rowlive) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {                 // candidate rows d0-1 .. d0+1 = region rows ky+1 .. ky+3
            const uint32_t* rr = rwords + (ky + dy + 1) * (RSB / 4);
            const uint32_t w0 = rr[0], w1 = rr[1], w2 = rr[2];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {             // candidate cols d1-1 .. d1+1 = region cols +1 .. +3
                const uint32_t b0 = __funnelshift_r(w0, w1, 8 * (dx + 1));
                const uint32_t b1 = __funnelshift_r(w1, w2, 8 * (dx + 1));
                acc[dy * 3 + dx] = __dp4a(a1, b1, __dp4a(a0, b0, acc[dy * 3 + dx]));
            }
        }
    }
    // butterfly reduction of 16 partial sums in 16 shuffles: afterwards this lane holds the
    // warp total of candidate my_s = b4*8 + b3*4 + b2*2 + b1 (b_k = bit k of the lane id)
    uint32_t part = 0;
    {
        uint32_t w8[8], w4[4], w2[2];
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
        for (int q = 0; q < 8; ++q) w8[q] = (h16 ? acc[q + 8] : acc[q]) + __shfl_xor_sync(0xffffffffu, h16 ? acc[q] : acc[q + 8], 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) w4[q] = (h8 ? w8[q + 4] : w8[q]) + __shfl_xor_sync(0xffffffffu, h8 ? w8[q] : w8[q + 4], 8);
#pragma unroll
        for (int q = 0; q < 2; ++q) w2[q] = (h4 ? w4[q + 2] : w4[q]) + __shfl_xor_sync(0xffffffffu, h4 ? w4[q] : w4[q + 2], 4);
        part = (h2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? w2[0] : w2[1], 2);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
    }
    const int my_s = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    float myval = 0.0f;
    {
        const int qy = d0 + my_s / 3 - 1, qx = d1 + my_s % 3 - 1;
        if (my_s < 9 && qy >= 0 && qy < T0 && qx >= 0 && qx < T1) myval = value_at((int)part, qy, qx);
    }
    float best = 0.f, centre = 0.f;
    int bi = 0;
    bool best_nan = false;
#pragma unroll
    for (int s = 0; s < 9; ++s) {
        const int src = ((s >> 3) & 1) * 16 + ((s >> 2) & 1) * 8 + ((s >> 1) & 1) * 4 + (s & 1) * 2;
        const float v = __shfl_sync(0xffffffffu, myval, src);
        if (s == 4) centre = v;
        if (s == 0) { best = v; best_nan = (v != v); }
        else if (!best_nan && (v > best || v != v)) { best = v; bi = s; best_nan = (v != v); }
    }
    if (best < DM_NEAR_ZERO_F) { bi = 4; best = centre; }
    const int c0 = d0 + bi / 3 - 1, c1 = d1 + bi % 3 - 1;
    const float score = best + centre;

    // ---- misc/Matching.py:165-209 parabola fit (index -1 wraps, upper edge skipped)
    double mrow = (double)c0, mcol = (double)c1;
    if (a.sub_pix && c0 >= 0 && c0 < T0 && c1 >= 0 && c1 < T1) {
        // neighbours: 0 = (c0+1,c1)  1 = (c0-1 | wrap, c1)  2 = (c0,c1+1)  3 = (c0, c1-1 | wrap)
        const int ny[4] = {c0 + 1, (c0 == 0 ? T0 - 1 : c0 - 1), c0, c0};
        const int nx[4] = {c1, c1, c1 + 1, (c1 == 0 ? T1 - 1 : c1 - 1)};
        const bool nok[4] = {c0 + 1 < T0, c0 + 1 < T0, c1 + 1 < T1, c1 + 1 < T1};
        uint32_t nacc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            nacc[q] = 0;
            if (!nok[q] || !rowlive) continue;
            const int ry = ny[q] - (d0 - 2), rx = nx[q] - (d1 - 2);
            if (ry >= 0 && ry <= 4 && rx >= 0 && rx <= 4) {         // inside the staged region (warp-uniform)
                const int bo = hf * 8 + rx;
                const uint32_t* rr = reinterpret_cast<const uint32_t*>(region) + (ky + ry) * (RSB / 4) + (bo >> 2);
                const uint32_t w0 = rr[0], w1 = rr[1], w2 = rr[2];
                const int sh = (bo & 3) * 8;
                nacc[q] = __dp4a(a1, __funnelshift_r(w1, w2, sh), __dp4a(a0, __funnelshift_r(w0, w1, sh), 0u));
            } else {                                                // wrapped index: far away, read global memory
                const uint8_t* b = a.img2 + (size_t)(oy + ny[q] + ky) * a.pitch + ox + nx[q] + hf * 8;
                uint32_t b0 = 0, b1 = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t v = (hf * 8 + u < WS) ? (uint32_t)__ldg(b + u) : 0u;
                    if (u < 4) b0 |= v << (8 * u); else b1 |= v << (8 * (u - 4));
                }
                nacc[q] = __dp4a(a1, b1, __dp4a(a0, b0, 0u));
            }
        }
        // 4 partial sums -> lane holds the total of neighbour my_n = b4*2 + b3 (6 shuffles)
        const bool h16 = lane & 16, h8 = lane & 8;
        uint32_t u0 = (h16 ? nacc[2] : nacc[0]) + __shfl_xor_sync(0xffffffffu, h16 ? nacc[0] : nacc[2], 16);
        uint32_t u1 = (h16 ? nacc[3] : nacc[1]) + __shfl_xor_sync(0xffffffffu, h16 ? nacc[1] : nacc[3], 16);
        uint32_t tot = (h8 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, h8 ? u0 : u1, 8);
        tot += __shfl_xor_sync(0xffffffffu, tot, 4);
        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        const int my_n = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
        const int sy = my_n == 0 ? ny[0] : (my_n == 1 ? ny[1] : c0);
        const int sx = my_n == 2 ? nx[2] : (my_n == 3 ? nx[3] : c1);
        const bool sok = my_n < 2 ? nok[0] : nok[2];
        float nval = 0.0f;
        if (sok) nval = value_at((int)tot, sy, sx);
        const float r0 = best;                      // level-0 value at the match itself
        const float v0 = __shfl_sync(0xffffffffu, nval, 0), v1 = __shfl_sync(0xffffffffu, nval, 8);
        const float v2 = __shfl_sync(0xffffffffu, nval, 16), v3 = __shfl_sync(0xffffffffu, nval, 24);
        if (nok[0] && r0 > v0 && r0 > v1) mrow += (double)(-(v0 - v1) / (2.0f * (v0 + v1 - 2.0f * r0)));
        if (nok[2] && r0 > v2 && r0 > v3) mcol += (double)(-(v2 - v3) / (2.0f * (v2 + v3 - 2.0f * r0)));
    }
    if (lane != 0) return;

    // ---- planes + paste (misc/Calc_difference.py:36-48, misc/image_cut_solver.py:165-175)
    const int g = a.first_tile + n;
    const int gi = g / a.len1, gj = g - gi * a.len1;
    const int Y = a.s0 * gi + i, X = a.s1 * gj + j;
    if (min(Y / a.s0, a.len0 - 1) != gi || min(X / a.s1, a.len1 - 1) != gj) return;   // a later tile owns this pixel
    const double e0 = __dsub_rn((double)i, mrow), e1 = __dsub_rn((double)j, mcol);
    const size_t plane = (size_t)a.out_h * a.out_w, pix = (size_t)Y * a.out_w + X;
    for (int m = 0; m < a.n_modes; ++m) {
        const double v = a.modes[m] == DM_MODE_ELEVATION ? e1
                       : a.modes[m] == DM_MODE_ELEVATION2 ? e0
                       : __dsqrt_rn(__fma_rn(e1, e1, __dmul_rn(e0, e0)));
        a.d_map[m * plane + pix] = v;
    }
    a.out_map[pix] = (double)score;
}

template <int WS>
static void launch_final(const FinalArgs& fa, long long n_patches, cudaStream_t st) {
    dm_final_level_kernel<WS><<<dm_div_up(n_patches, 8), 256, 0, st>>>(fa, n_patches);
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
        dm_tile_origin_kernel2<<<dm_div_up(nt, 128), 128, 0, st>>>(fb.origin, nt, a->first_tile, a->len1, a->s0, a->s1);
        DM_LAUNCH_CHECK();
        if ((rc = dm_descriptors(a->img1, a->scene_h, a->scene_w, a->scene_w, fb.origin, nt, t0, t1, a->ws, fb.desc1, fb.stat1, st)) != DM_OK) return rc;
        if ((rc = dm_descriptors(a->img2, a->scene_h, a->scene_w, a->scene_w, fb.origin, nt, t0, t1, a->ws, fb.desc2, fb.stat2, st)) != DM_OK) return rc;
        ctx->launches[DM_STAGE_DESCRIPTORS] += 3;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    {
        StageTimer tm(ctx, DM_STAGE_CORRELATION);
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        if ((rc = dm_correlation_umma_pool(fb.desc1, fb.stat1, fb.desc2, fb.stat2, nt, t0, t1, a->kpad, a->method,
                                           fb.pooled, fb.rowmin, fb.rowmax, st)) != DM_OK) return rc;
        ctx->launches[DM_STAGE_CORRELATION] += 1;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    {
        StageTimer tm(ctx, DM_STAGE_NORMALIZE);        // min-max + rectify + first child average
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        const long long parents = (long long)nt * (P / 4);
        const int q4 = P / 16;
        const int threads = q4 >= 256 ? 256 : (q4 < 32 ? 32 : q4);
        dm_aggregate_first_kernel<<<(unsigned)parents, threads, 0, st>>>(fb.pooled, fb.rowmin, fb.rowmax, t0, t1, fb.level[1]);
        DM_LAUNCH_CHECK();
        ctx->launches[DM_STAGE_NORMALIZE] += 1;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    {
        StageTimer tm(ctx, DM_STAGE_AGGREGATE);
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        for (int k = 1; k + 1 < L; ++k) {
            if ((rc = dm_aggregate(fb.level[k], nt, t0 >> k, t1 >> k, t0 >> k, t1 >> k, 1, fb.level[k + 1], st)) != DM_OK) return rc;
            ctx->launches[DM_STAGE_AGGREGATE] += 1;
        }
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    int cur = 0;
    {
        StageTimer tm(ctx, DM_STAGE_BACKTRACK);
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        if ((rc = dm_backtrack_top(fb.level[L - 1], 0, nt, t0 >> (L - 1), t1 >> (L - 1), fb.match[cur], fb.score[cur], st)) != DM_OK) return rc;
        ctx->launches[DM_STAGE_BACKTRACK] += 1;
        for (int k = L - 2; k >= 1; --k) {
            if ((rc = dm_backtrack_level(fb.level[k], 0, nt, t0 >> k, t1 >> k, t0 >> k, t1 >> k, fb.match[cur], fb.match[cur ^ 1], fb.score[cur ^ 1], st)) != DM_OK) return rc;
            cur ^= 1;
            ctx->launches[DM_STAGE_BACKTRACK] += 1;
        }
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    {
        StageTimer tm(ctx, DM_STAGE_PLANES);           // level-0 backtracking + sub-pixel + planes
        if ((rc = tm.begin(ck)) != DM_OK) return rc;
        FinalArgs fa;
        fa.img1 = a->img1; fa.img2 = a->img2; fa.pitch = a->scene_w;
        fa.origin = fb.origin; fa.stat1 = (const dm_stat*)fb.stat1; fa.stat2 = (const dm_stat*)fb.stat2;
        fa.rowmin = fb.rowmin; fa.rowmax = fb.rowmax; fa.parent = fb.match[cur];
        fa.t0 = t0; fa.t1 = t1; fa.ws = a->ws; fa.normed = a->method == DM_TM_CCOEFF_NORMED; fa.sub_pix = a->sub_pix;
        fa.n_modes = a->n_modes; for (int m = 0; m < 4; ++m) fa.modes[m] = a->modes[m];
        fa.s0 = a->s0; fa.s1 = a->s1; fa.len0 = a->len0; fa.len1 = a->len1; fa.out_h = a->out_h; fa.out_w = a->out_w;
        fa.first_tile = a->first_tile; fa.d_map = a->d_map; fa.out_map = a->out_map;
        const long long n_patches = (long long)nt * P;
        fa.scene_h = a->scene_h;
        switch (a->ws) {
            case 3: launch_final<3>(fa, n_patches, st); break;
            case 5: launch_final<5>(fa, n_patches, st); break;
            case 7: launch_final<7>(fa, n_patches, st); break;
            case 9: launch_final<9>(fa, n_patches, st); break;
            case 11: launch_final<11>(fa, n_patches, st); break;
            case 13: launch_final<13>(fa, n_patches, st); break;
            case 15: launch_final<15>(fa, n_patches, st); break;
            default: DM_REQUIRE(false, DM_ERR_UNSUPPORTED, "fused path supports odd window sizes 3..15 (got %d)", a->ws);
        }
        DM_LAUNCH_CHECK();
        ctx->launches[DM_STAGE_PLANES] += 1;
        if ((rc = tm.end()) != DM_OK) return rc;
    }
    return DM_OK;
}
