// Patch descriptors: replaces Correlation_map._create_atomic_patch
// (misc/Correlation_map.py:51-67) and the window statistics cv2.matchTemplate derives
// from integral images (misc/Feature_value.py:41).
//
// One warp per patch.  The ws*ws window is centred on the patch's own rounded mean so
// every value is an integer in [-255,255] -- exactly representable in bf16 -- and the
// residual sum S' is bounded by K/2, which keeps the fp32 ZNCC numerator
// dot - S1'*S2'/K free of cancellation.  Row layout [n][P][kpad], K-major, zero padded:
// the layout TMA loads straight into 128B-swizzled shared memory for tcgen05.mma.
#include <cstdlib>

#include "dm_common.cuh"
#include "dm_internal.h"

// K layout of a descriptor row: k = ky * rstride + kx with rstride = 8 / 16 / 32 (window rows
// padded to a power of two, so a lane owns 8 consecutive entries of ONE window row); both
// images use the same order, the contraction does not care.
static __host__ __device__ inline int dm_row_stride(int ws) { return ws <= 8 ? 8 : (ws <= 16 ? 16 : 32); }

int dm_desc_kreal(int ws) { return ws * dm_row_stride(ws); }

// Correction slots.  ZNCC needs  dot - S1' * S2' / K  (S' = residual sum of a centred window).  A window row
// occupies ws of its rstride entries, so every descriptor row has unused entries; three of them carry the
// correction INTO the contraction: the patch side (side 1) stores S1' three times (an integer, |S'| <= K/2,
// exact in bf16), the search side (side 2) stores the three bf16 parts hi + mid + lo of the float32 value
// -S2'/K (an exact split: 3 x 8 mantissa bits).  The tensor core then accumulates
//     sum a1' a2'  +  S1' (hi + mid + lo)  =  dot - S1' * fl(S2'/K)
// -- every product is exact, only the accumulation of the three fractional terms rounds (<= 1 ulp of the
// accumulator each) -- and the epilogue of the tcgen05 kernel is left with ONE per-column factor (inv2)
// instead of two: half the column-table traffic and one FMA per element less.
// The slots are the first three unused entries in (ky, kx) order: kx in [ws, rstride).
__host__ __device__ inline void dm_desc_slots(int ws, int k[3]) {
    const int r = dm_row_stride(ws);
    int n = 0;
    for (int ky = 0; ky < ws && n < 3; ++ky)
        for (int kx = ws; kx < r && n < 3; ++kx) k[n++] = ky * r + kx;
}
void dm_desc_special_slots(int ws, int k[3]) { dm_desc_slots(ws, k); }

// the three bf16 parts of a float32 value (hi + mid + lo == v exactly for |v| in the normal range)
__device__ __forceinline__ void dm_split3(float v, float part[3]) {
    const float hi = __bfloat162float(__float2bfloat16_rn(v));
    const float r1 = __fsub_rn(v, hi);
    const float mid = __bfloat162float(__float2bfloat16_rn(r1));
    const float r2 = __fsub_rn(r1, mid);
    part[0] = hi; part[1] = mid; part[2] = __bfloat162float(__float2bfloat16_rn(r2));
}
// value of correction slot j for a window with residual sum rs: side 1 -> S', side 2 -> part j of -S'/K
__device__ __forceinline__ float dm_slot_value(int side, int j, int rs, int K) {
    if (side == 1) return (float)rs;
    if (side != 2) return 0.0f;
    float part[3];
    dm_split3(-__fdiv_rn((float)rs, (float)K), part);
    return part[j];
}

extern "C" int dm_kpad(int ws) {
    int k = ws * dm_row_stride(ws);
    return ((k + 63) / 64) * 64;       // multiple of 64 bf16 = one 128-byte swizzle row
}

__device__ __forceinline__ void dm_write_stats(dm_stat* stat, long long n_patches, long long pidx, int K, int S, int Q, int mean) {
    // S, Q = sum and sum of squares of the raw window; residuals after centring on `mean`
    const int rs = S - K * mean;                                  // S'
    const int rq = Q - 2 * mean * S + K * mean * mean;            // sum (a - mean)^2
    const float s = (float)rs;
    const float sk = __fdiv_rn(s, (float)K);
    const float var = __fsub_rn((float)rq, __fmul_rn(s, sk));     // sum a'^2 - S'^2/K
    const bool flat = (rq == 0);
    const float inv = flat ? 0.0f : __frcp_rn(__fsqrt_rn(var));
    stat[pidx] = make_float4(s, inv, sk, (float)mean);
    // compact column table of the tcgen05 epilogue: inv alone (the S'/K term rides in the correction slots)
    reinterpret_cast<float*>(stat + n_patches)[pidx] = inv;
}

// Generic kernel (any odd ws <= 31, any grid): one warp per patch, two passes over the window.
__global__ void __launch_bounds__(256)
dm_descriptor_kernel(const uint8_t* __restrict__ scene, int pitch,
                     const int32_t* __restrict__ origin_yx, long long n_patches,
                     int t0, int t1, int ws, int kpad, int side,
                     __nv_bfloat16* __restrict__ desc, dm_stat* __restrict__ stat) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n_patches) return;
    const int P = t0 * t1;
    const int tile = (int)(warp / P);
    const int p = (int)(warp - (long long)tile * P);
    const int i = p / t1, j = p - i * t1;
    const int K = ws * ws, rstr = dm_row_stride(ws);
    const uint8_t* base = scene + (size_t)(origin_yx[2 * tile] + i) * pitch + origin_yx[2 * tile + 1] + j;

    int sum = 0, sq = 0;
    for (int k = lane; k < K; k += 32) {
        int ky = k / ws, kx = k - ky * ws;
        int v = base[ky * pitch + kx];
        sum += v; sq += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    const int mean = dm_round_mean(sum, K);
    __nv_bfloat16* row = desc + (size_t)warp * kpad;
    int slot[3];
    dm_desc_slots(ws, slot);
    const int rs = sum - K * mean;
    for (int k = lane; k < kpad; k += 32) {
        const int ky = k / rstr, kx = k - ky * rstr;
        float v = 0.f;
        if (ky < ws && kx < ws) v = (float)((int)base[ky * pitch + kx] - mean);
        for (int j = 0; j < 3; ++j) if (k == slot[j]) v = dm_slot_value(side, j, rs, K);
        row[k] = __float2bfloat16(v);
    }
    if (lane == 0) dm_write_stats(stat, n_patches, warp, K, sum, sq, mean);
}

// Fast kernel for ws <= 15 and T1 % 8 == 0.  A group of lanes produces the descriptors of 8
// consecutive patches of one grid row: lane (ky, hf) reads the 15 scene bytes of window row
// ky that those 8 windows share, keeps them packed in 4 words, and slides over them with
// funnel shifts.  Sums come from DP4A + REDUX, the centred values are formed with the
// 0x4B000000 byte-to-float trick, and every lane writes its 8 entries of each patch as one
// 16-byte store (a descriptor row leaves the group as one contiguous segment).
template <int WS>
__global__ void __launch_bounds__(256)
dm_descriptor_row_kernel(const uint8_t* __restrict__ scene, int pitch,
                         const int32_t* __restrict__ origin_yx, long long n_groups, long long n_patches,
                         int t0, int t1, dm_fastdiv fd_jb, dm_fastdiv fd_t0, int side,
                         __nv_bfloat16* __restrict__ desc, dm_stat* __restrict__ stat,
                         const uint8_t* __restrict__ scene_b, __nv_bfloat16* __restrict__ desc_b, dm_stat* __restrict__ stat_b) {
    dm_pdl_wait();                  // the tile origins come from the kernel in front
    dm_pdl_launch_dependents();
    // both images of a pair in one launch (dm_descriptors_both): grid row 1 is the search image
    if (blockIdx.y) { scene = scene_b; desc = desc_b; stat = stat_b; side = 2; }
    constexpr int K = WS * WS;
    constexpr int RSTR = WS <= 8 ? 8 : 16;
    constexpr int KPAD = ((WS * RSTR + 63) / 64) * 64;
    constexpr int LPG = RSTR == 16 ? 32 : 8;                          // lanes per group
    // correction slots (dm_desc_slots): S' is an integer in [-K/2, K/2], so the three bf16 parts of -S'/K
    // (side 2) come from a table of K + 1 entries built once per CTA
    // Small windows (rstride 8: one lane per window row, the three slots in one lane) take the slot values
    // inside the row stores; large ones (two lanes per row, the slots in three lanes) are cheaper with three
    // 2-byte stores per patch after the loop (measured: ws 5 1.14 -> 0.94 ms on 64 x 512^2, ws 15 0.187 vs 0.208 ms).
    constexpr int NF = RSTR - WS;                                     // unused entries per window row
    constexpr bool SLOT_IN_LOOP = RSTR == 8;
    __shared__ float slot_tab[3][SLOT_IN_LOOP ? K / 2 * 2 + 2 : 1];
    if (SLOT_IN_LOOP && side == 2) {
        for (int idx = threadIdx.x; idx <= K / 2 * 2; idx += blockDim.x) {
            float part[3];
            dm_split3(-__fdiv_rn((float)(idx - K / 2), (float)K), part);
            slot_tab[0][idx] = part[0]; slot_tab[1][idx] = part[1]; slot_tab[2][idx] = part[2];
        }
        __syncthreads();
    }
    constexpr int GPW = 32 / LPG;                                     // groups per warp
    const int lane = threadIdx.x & 31, gl = lane % LPG;
    const unsigned gmask = LPG == 32 ? 0xffffffffu : (0xffu << (lane & ~7));
    long long g = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPG;
    const bool live = g < n_groups;
    if (!live) g = n_groups - 1;                                      // keep the group converged for REDUX
    // group -> (tile, grid row i, block of 8 columns jb) with exact multiply-shift divisions (the
    // group count fits 32 bits: the host checks it); three 64-bit `/` and `%` here were a quarter
    // of the kernel's instructions
    const int jb_n = t1 >> 3;
    const uint32_t g32 = (uint32_t)g;
    const uint32_t t = dm_fd_div(g32, fd_jb);
    const int jb = (int)(g32 - t * (uint32_t)jb_n);
    const int tile = (int)dm_fd_div(t, fd_t0);
    const int i = (int)(t - (uint32_t)tile * (uint32_t)t0);
    const int j0 = jb * 8;
    const int ky = RSTR == 16 ? (gl >> 1) : gl, hf = RSTR == 16 ? (gl & 1) : 0;
    const bool rowlive = ky < WS;
    // bytes of scene row (i + ky) this lane needs: columns j0 + 8*hf + [0, nb)
    const int nb = !rowlive ? 0 : (hf == 0 ? (7 + WS < 15 ? 7 + WS : 15) : WS - 8 + 7);
    const uint8_t* src = scene + (size_t)(origin_yx[2 * tile] + i + (rowlive ? ky : 0)) * pitch + origin_yx[2 * tile + 1] + j0 + hf * 8;
    // the (up to 15) bytes as aligned 32-bit loads + funnel shifts: 5 requests instead of 15 (the
    // kernel was L1/TEX-bound on the byte loads).  Only words that hold a needed byte are read,
    // so nothing outside the scene's own 4-byte-aligned words is touched.
    uint32_t w[4];
    {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
        const uint32_t* ap = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
        const int off = (int)(addr & 3);
        uint32_t a[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) a[k] = (4 * k < off + nb) ? __ldg(ap + k) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int nvk = nb - 4 * k;                 // valid bytes of word k
            const uint32_t m = nvk >= 4 ? 0xffffffffu : (nvk <= 0 ? 0u : ((1u << (8 * nvk)) - 1u));
            w[k] = __funnelshift_r(a[k], a[k + 1], 8 * off) & m;
        }
    }
    // valid bytes of an 8-byte window slice: window columns 8*hf + u < WS
    const int nv = !rowlive ? 0 : (WS - hf * 8 > 8 ? 8 : WS - hf * 8);
    const uint32_t mlo = nv >= 4 ? 0xffffffffu : ((1u << (8 * nv)) - 1u);
    const uint32_t mhi = nv <= 4 ? 0u : (nv >= 8 ? 0xffffffffu : ((1u << (8 * (nv - 4))) - 1u));
    const long long p0 = ((long long)tile * t0 + i) * t1 + j0;        // first patch of the group
    int myS = 0, myQ = 0, myMean = 0;                                 // lane gl < 8 keeps the sums of patch gl
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t lo = __funnelshift_r(w[q >> 2], w[(q >> 2) + 1], 8 * (q & 3)) & mlo;
        const uint32_t hi = __funnelshift_r(w[(q >> 2) + 1], w[(q >> 2) + 2], 8 * (q & 3)) & mhi;
        const int s = (int)__dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, 0u));
        const int sq = (int)__dp4a(lo, lo, __dp4a(hi, hi, 0u));
        // window sums over the lanes of the group.  A whole warp (two lanes per window row) uses REDUX; the four
        // 8-lane groups of a warp would each bring their own member mask, which the compiler serialises group by
        // group behind WARPSYNC / ENDCOLLECTIVE (47 convergence regions in the ws = 5 kernel) -- three full-warp
        // butterfly steps stay inside the aligned groups of 8 and cost six shuffles
        int S = s, Q = sq;
        if (LPG == 32) { S = __reduce_add_sync(gmask, s); Q = __reduce_add_sync(gmask, sq); }
        else {
#pragma unroll
            for (int o = 1; o < LPG; o <<= 1) { S += __shfl_xor_sync(0xffffffffu, S, o); Q += __shfl_xor_sync(0xffffffffu, Q, o); }
        }
        const int mean = dm_round_mean(S, K);
        if (!live) continue;
        if (gl * 8 < KPAD) {
            // invalid positions take the value `mean` so that they centre to exactly 0
            const uint32_t mw = (uint32_t)mean * 0x01010101u;
            const uint32_t l2 = lo | (mw & ~mlo), h2 = hi | (mw & ~mhi);
            const float off = 8388608.0f + (float)mean;              // 2^23 + mean
            float f[8];
            f[0] = __uint_as_float(__byte_perm(l2, 0x4B000000u, 0x7650)) - off;
            f[1] = __uint_as_float(__byte_perm(l2, 0x4B000000u, 0x7651)) - off;
            f[2] = __uint_as_float(__byte_perm(l2, 0x4B000000u, 0x7652)) - off;
            f[3] = __uint_as_float(__byte_perm(l2, 0x4B000000u, 0x7653)) - off;
            f[4] = __uint_as_float(__byte_perm(h2, 0x4B000000u, 0x7650)) - off;
            f[5] = __uint_as_float(__byte_perm(h2, 0x4B000000u, 0x7651)) - off;
            f[6] = __uint_as_float(__byte_perm(h2, 0x4B000000u, 0x7652)) - off;
            f[7] = __uint_as_float(__byte_perm(h2, 0x4B000000u, 0x7653)) - off;
            if (SLOT_IN_LOOP && side != 0) {
                const int rs = S - K * mean;                        // |rs| <= K/2
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    constexpr int dummy = 0; (void)dummy;
                    const int sky = j / NF, skx = WS + j % NF;      // the j-th unused entry in (ky, kx) order
                    if (ky == sky && hf == skx / 8) {
                        const float sv = side == 1 ? (float)rs : slot_tab[j][rs + K / 2];
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (u == skx % 8) f[u] = sv;
                    }
                }
            }
            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&b0); pk.y = *reinterpret_cast<uint32_t*>(&b1);
            pk.z = *reinterpret_cast<uint32_t*>(&b2); pk.w = *reinterpret_cast<uint32_t*>(&b3);
            *reinterpret_cast<uint4*>(desc + (size_t)(p0 + q) * KPAD + gl * 8) = pk;
        }
        if (gl == q) { myS = S; myQ = Q; myMean = mean; }
    }
    // the statistics (a division, a square root and a reciprocal, ~60 instructions) once for the
    // eight patches in parallel instead of once per patch behind a one-lane branch
    if (!SLOT_IN_LOOP) __syncwarp();    // the slot stores below overwrite zeros other lanes of the group have just stored
    if (live && gl < 8) {
        dm_write_stats(stat, n_patches, p0 + gl, K, myS, myQ, myMean);
        if (!SLOT_IN_LOOP && side != 0) {
            const int rs = myS - K * myMean;
            float part[3] = {(float)rs, (float)rs, (float)rs};
            if (side == 2) dm_split3(-__fdiv_rn((float)rs, (float)K), part);
            __nv_bfloat16* row = desc + (size_t)(p0 + gl) * KPAD;
#pragma unroll
            for (int j = 0; j < 3; ++j) row[(j / NF) * RSTR + WS + j % NF] = __float2bfloat16_rn(part[j]);
        }
    }
}

template <int WS>
static void launch_descriptor_row(const uint8_t* scene, int pitch, const int32_t* origin, long long n_patches,
                                  int t0, int t1, int side, void* desc, float* stat, cudaStream_t st,
                                  const uint8_t* scene_b = nullptr, void* desc_b = nullptr, float* stat_b = nullptr) {
    constexpr int RSTR = WS <= 8 ? 8 : 16;
    constexpr int GPW = RSTR == 16 ? 1 : 4;
    const long long n_groups = n_patches / 8;
    const long long warps = (n_groups + GPW - 1) / GPW;
    dim3 grid((unsigned)dm_div_up(warps, 8), scene_b ? 2 : 1);
    dm_launch_dep(DM_PDL_DESC, dm_descriptor_row_kernel<WS>, grid, dim3(256), 0, st, scene, pitch, origin, n_groups, n_patches, t0, t1,
                  dm_make_fastdiv((uint32_t)(t1 >> 3)), dm_make_fastdiv((uint32_t)t0), side,
                  (__nv_bfloat16*)desc, (dm_stat*)stat, scene_b, (__nv_bfloat16*)desc_b, (dm_stat*)stat_b);
}

// Descriptors of both images of a pair (side 1 = patch image, side 2 = search image) over the same tile origins:
// one launch with a two-row grid where the row kernel applies, else two launches.  *launches = kernels launched.
int dm_descriptors_both(const uint8_t* img1, const uint8_t* img2, int scene_h, int scene_w, int pitch,
                        const int32_t* origin_yx_dev, int n_tiles, int t0, int t1, int ws,
                        void* desc1, float* stat1, void* desc2, float* stat2, cudaStream_t st, int* launches) {
    static const bool split = getenv("DM_DESC_SPLIT") != nullptr;        // measurement aid: one launch per image
    const long long n_patches = (long long)n_tiles * t0 * t1;
    const bool rowk = (t1 % 8 == 0) && n_patches / 8 < (1LL << 32) && ws >= 3 && ws <= 15 && (ws & 1);
    const bool sized = scene_h >= t0 + ws - 1 && scene_w >= t1 + ws - 1 && pitch >= scene_w && t0 > 0 && t1 > 0 && n_tiles > 0;
    if (split || !rowk || !sized) {
        int rc = dm_descriptors(img1, scene_h, scene_w, pitch, origin_yx_dev, n_tiles, t0, t1, ws, 1, desc1, stat1, st);
        if (rc != DM_OK) return rc;
        if (launches) *launches = 2;
        return dm_descriptors(img2, scene_h, scene_w, pitch, origin_yx_dev, n_tiles, t0, t1, ws, 2, desc2, stat2, st);
    }
    switch (ws) {
        case 3:  launch_descriptor_row<3>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
        case 5:  launch_descriptor_row<5>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
        case 7:  launch_descriptor_row<7>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
        case 9:  launch_descriptor_row<9>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
        case 11: launch_descriptor_row<11>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
        case 13: launch_descriptor_row<13>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
        default: launch_descriptor_row<15>(img1, pitch, origin_yx_dev, n_patches, t0, t1, 1, desc1, stat1, st, img2, desc2, stat2); break;
    }
    DM_LAUNCH_CHECK();
    if (launches) *launches = 1;
    return DM_OK;
}

extern "C" int dm_descriptors(const uint8_t* scene_dev, int scene_h, int scene_w, int pitch,
                              const int32_t* origin_yx_dev, int n_tiles, int t0, int t1, int ws, int side,
                              void* desc_bf16_dev, float* stat_dev, void* stream) {
    DM_REQUIRE(side >= 0 && side <= 2, DM_ERR_INVALID, "dm_descriptors: side must be 1 (patch image), 2 (search image) or 0 (got %d)", side);
    DM_REQUIRE(ws >= 1 && (ws & 1) && ws <= 31, DM_ERR_INVALID, "dm_descriptors: window_size must be odd and <= 31 (got %d)", ws);
    DM_REQUIRE(t0 > 0 && t1 > 0 && n_tiles > 0, DM_ERR_INVALID, "dm_descriptors: empty grid");
    DM_REQUIRE(scene_h >= t0 + ws - 1 && scene_w >= t1 + ws - 1 && pitch >= scene_w, DM_ERR_INVALID,
               "dm_descriptors: scene %dx%d smaller than a tile", scene_h, scene_w);
    const long long n_patches = (long long)n_tiles * t0 * t1;
    cudaStream_t st = (cudaStream_t)stream;
    const bool rowk = (t1 % 8 == 0) && n_patches / 8 < (1LL << 32);
    switch (rowk ? ws : 0) {
        case 3:  launch_descriptor_row<3>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        case 5:  launch_descriptor_row<5>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        case 7:  launch_descriptor_row<7>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        case 9:  launch_descriptor_row<9>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        case 11: launch_descriptor_row<11>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        case 13: launch_descriptor_row<13>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        case 15: launch_descriptor_row<15>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, side, desc_bf16_dev, stat_dev, st); break;
        default: {
            const int warps = 8;
            dm_descriptor_kernel<<<dm_div_up(n_patches, warps), warps * 32, 0, st>>>(
                scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, ws, dm_kpad(ws), side,
                (__nv_bfloat16*)desc_bf16_dev, (dm_stat*)stat_dev);
        }
    }
    DM_LAUNCH_CHECK();
    return DM_OK;
}

// ---------------------------------------------------------------------------------------
// Feature_value.__call__ for arbitrary sizes (misc/Feature_value.py:39-43): one thread
// per output position, exact 64-bit integer sums, OpenCV's float64 normalisation
// (templmatch.cpp::common_matchTemplate) restated, float32 result.  Not a hot path --
// it exists so that callers such as for_igarss/cor_map.py:33-35 (49x49 patch on a
// 500x500 image) keep working.
// ---------------------------------------------------------------------------------------
__global__ void dm_feature_value_kernel(const uint8_t* __restrict__ patch, int ph, int pw,
                                        const uint8_t* __restrict__ image, int ih, int iw,
                                        int normed, float* __restrict__ out) {
    const int ow = iw - pw + 1, oh = ih - ph + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ow * oh) return;
    const int y = idx / ow, x = idx - y * ow;
    long long cc = 0, s1 = 0, q1 = 0, s2 = 0, q2 = 0;
    for (int ky = 0; ky < ph; ++ky) {
        const uint8_t* ir = image + (size_t)(y + ky) * iw + x;
        const uint8_t* pr = patch + (size_t)ky * pw;
        for (int kx = 0; kx < pw; ++kx) {
            int a = pr[kx], b = ir[kx];
            cc += a * b; s1 += a; q1 += a * a; s2 += b; q2 += b * b;
        }
    }
    const double inv_area = 1.0 / ((double)ph * pw);
    const double templ_mean = (double)s1 * inv_area;
    double num = (double)cc - templ_mean * (double)s2;
    if (normed) {
        double tv = fmax((double)q1 * inv_area - templ_mean * templ_mean, 0.0);
        double sd = sqrt(tv);
        double tn2 = sd * sd;
        if (tn2 < 2.220446049250313e-16) {
            num = 1.0;
        } else {
            double templ_norm = sqrt(tn2) / sqrt(inv_area);
            double wsum2 = (double)q2;
            double diff2 = fmax(wsum2 - (double)s2 * (double)s2 * inv_area, 0.0);
            double t = (diff2 <= fmin(0.5, 10.0 * 1.1920928955078125e-07 * wsum2)) ? 0.0 : sqrt(diff2) * templ_norm;
            if (fabs(num) < t) num /= t;
            else if (fabs(num) < t * 1.125) num = num > 0 ? 1.0 : -1.0;
            else num = 0.0;
        }
    }
    out[idx] = (float)num;
}

int dm_minmax_rectify_impl(const float* raw, long long rows, int q, float* norm, float* rect,
                           float* rowmin, float* rowmax, cudaStream_t stream);

extern "C" int dm_feature_value(const uint8_t* patch_dev, int ph, int pw,
                                const uint8_t* image_dev, int ih, int iw,
                                int method, float* out_dev, void* stream) {
    DM_REQUIRE(method == DM_TM_CCOEFF || method == DM_TM_CCOEFF_NORMED, DM_ERR_INVALID,
               "dm_feature_value: invalid method %d", method);
    DM_REQUIRE(ph >= 1 && pw >= 1 && ih >= ph && iw >= pw, DM_ERR_INVALID,
               "dm_feature_value: patch %dx%d does not fit image %dx%d", ph, pw, ih, iw);
    const int n = (ih - ph + 1) * (iw - pw + 1);
    dm_feature_value_kernel<<<dm_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(
        patch_dev, ph, pw, image_dev, ih, iw, method == DM_TM_CCOEFF_NORMED, out_dev);
    DM_LAUNCH_CHECK();
    return dm_minmax_rectify_impl(out_dev, 1, n, out_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
