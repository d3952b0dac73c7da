// Patch descriptors: replaces Correlation_map._create_atomic_patch
// (misc/Correlation_map.py:51-67) and the window statistics cv2.matchTemplate derives
// from integral images (misc/Feature_value.py:41).
//
// One warp per patch.  The ws*ws window is centred on the patch's own rounded mean so
// every value is an integer in [-255,255] -- exactly representable in bf16 -- and the
// residual sum S' is bounded by K/2, which keeps the fp32 ZNCC numerator
// dot - S1'*S2'/K free of cancellation.  Row layout [n][P][kpad], K-major, zero padded:
// the layout TMA loads straight into 128B-swizzled shared memory for tcgen05.mma.
#include "dm_common.cuh"

extern "C" int dm_kpad(int ws) {
    int k = ws * ws;
    return ((k + 63) / 64) * 64;       // multiple of 64 bf16 = one 128-byte swizzle row
}

// Generic kernel (any odd ws <= 31): one warp per patch, two passes over the window.
__global__ void __launch_bounds__(256)
dm_descriptor_kernel(const uint8_t* __restrict__ scene, int pitch,
                     const int32_t* __restrict__ origin_yx, long long n_patches,
                     int t0, int t1, int ws, int kpad,
                     __nv_bfloat16* __restrict__ desc, dm_stat* __restrict__ stat) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n_patches) return;
    const int P = t0 * t1;
    const int tile = (int)(warp / P);
    const int p = (int)(warp - (long long)tile * P);
    const int i = p / t1, j = p - i * t1;
    const int K = ws * ws;
    const uint8_t* base = scene + (size_t)(origin_yx[2 * tile] + i) * pitch + origin_yx[2 * tile + 1] + j;

    int sum = 0;
    for (int k = lane; k < K; k += 32) {
        int ky = k / ws, kx = k - ky * ws;
        sum += base[ky * pitch + kx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int mean = dm_round_mean(sum, K);

    __nv_bfloat16* row = desc + (size_t)warp * kpad;
    int rs = 0, rq = 0;
    for (int k = lane; k < kpad; k += 32) {
        int v = 0;
        if (k < K) {
            int ky = k / ws, kx = k - ky * ws;
            v = (int)base[ky * pitch + kx] - mean;
        }
        rs += v;
        rq += v * v;
        row[k] = __float2bfloat16((float)v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
        rq += __shfl_xor_sync(0xffffffffu, rq, o);
    }
    if (lane == 0) {
        const float fk = (float)K;
        const float s = (float)rs;
        const float sk = __fdiv_rn(s, fk);
        const float var = __fsub_rn((float)rq, __fmul_rn(s, sk));     // sum a'^2 - S'^2/K
        const bool flat = (rq == 0);
        const float inv = flat ? 0.0f : __frcp_rn(__fsqrt_rn(var));
        stat[warp] = make_float4(s, inv, sk, (float)mean);
        // compact column table of the tcgen05 epilogue: per pair of patches {sk0, sk1, inv0, inv1}
        float* ct = reinterpret_cast<float*>(stat + n_patches) + (warp >> 1) * 4 + (warp & 1);
        ct[0] = sk; ct[2] = inv;
    }
}

// Fast kernel for the window sizes the repo uses (ws <= 15): the window is read once, every
// lane owns 8 consecutive K entries of its patch and writes them as one 16-byte store, so a
// descriptor row leaves the warp as coalesced 128..512-byte segments.  G = lanes per patch.
template <int WS>
__global__ void __launch_bounds__(256)
dm_descriptor_fast_kernel(const uint8_t* __restrict__ scene, int pitch,
                          const int32_t* __restrict__ origin_yx, long long n_patches, int t0, int t1,
                          __nv_bfloat16* __restrict__ desc, dm_stat* __restrict__ stat) {
    constexpr int K = WS * WS;
    constexpr int KPAD = ((K + 63) / 64) * 64;
    constexpr int LANES = KPAD / 8;                                   // lanes that store
    constexpr int G = LANES <= 8 ? 8 : (LANES <= 16 ? 16 : 32);       // lanes per patch
    constexpr int PER_WARP = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane % G;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    long long pidx = warp * PER_WARP + lane / G;
    const bool live = pidx < n_patches;
    if (!live) pidx = n_patches - 1;                                  // keep the warp converged for the shuffles
    const int P = t0 * t1;
    const int tile = (int)(pidx / P);
    const int p = (int)(pidx - (long long)tile * P);
    const int i = p / t1, j = p - i * t1;
    const uint8_t* base = scene + (size_t)(origin_yx[2 * tile] + i) * pitch + origin_yx[2 * tile + 1] + j;
    int px[8];
    int sum = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int k = gl * 8 + u;
        px[u] = 0;
        if (k < K) {
            const int ky = k / WS, kx = k - ky * WS;
            px[u] = base[ky * pitch + kx];
            sum += px[u];
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const int mean = dm_round_mean(sum, K);
    int rs = 0, rq = 0;
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int k = gl * 8 + u;
        const int v = (k < K) ? px[u] - mean : 0;
        rs += v;
        rq += v * v;
        f[u] = (float)v;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        rs += __shfl_xor_sync(0xffffffffu, rs, o);
        rq += __shfl_xor_sync(0xffffffffu, rq, o);
    }
    if (!live) return;
    if (gl < LANES) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(desc + (size_t)pidx * KPAD + gl * 8) = pk;
    }
    if (gl == 0) {
        const float s = (float)rs;
        const float sk = __fdiv_rn(s, (float)K);
        const float var = __fsub_rn((float)rq, __fmul_rn(s, sk));
        const bool flat = (rq == 0);
        const float inv = flat ? 0.0f : __frcp_rn(__fsqrt_rn(var));
        stat[pidx] = make_float4(s, inv, sk, (float)mean);
        float* ct = reinterpret_cast<float*>(stat + n_patches) + (pidx >> 1) * 4 + (pidx & 1);
        ct[0] = sk; ct[2] = inv;                                      // {sk0, sk1, inv0, inv1} per pair of patches
    }
}

template <int WS>
static void launch_descriptor_fast(const uint8_t* scene, int pitch, const int32_t* origin, long long n_patches,
                                   int t0, int t1, void* desc, float* stat, cudaStream_t st) {
    constexpr int K = WS * WS;
    constexpr int KPAD = ((K + 63) / 64) * 64;
    constexpr int LANES = KPAD / 8;
    constexpr int G = LANES <= 8 ? 8 : (LANES <= 16 ? 16 : 32);
    const long long warps = (n_patches + (32 / G) - 1) / (32 / G);
    dm_descriptor_fast_kernel<WS><<<dm_div_up(warps, 8), 256, 0, st>>>(scene, pitch, origin, n_patches, t0, t1,
                                                                        (__nv_bfloat16*)desc, (dm_stat*)stat);
}

extern "C" int dm_descriptors(const uint8_t* scene_dev, int scene_h, int scene_w, int pitch,
                              const int32_t* origin_yx_dev, int n_tiles, int t0, int t1, int ws,
                              void* desc_bf16_dev, float* stat_dev, void* stream) {
    DM_REQUIRE(ws >= 1 && (ws & 1) && ws <= 31, DM_ERR_INVALID, "dm_descriptors: window_size must be odd and <= 31 (got %d)", ws);
    DM_REQUIRE(t0 > 0 && t1 > 0 && n_tiles > 0, DM_ERR_INVALID, "dm_descriptors: empty grid");
    DM_REQUIRE(scene_h >= t0 + ws - 1 && scene_w >= t1 + ws - 1 && pitch >= scene_w, DM_ERR_INVALID,
               "dm_descriptors: scene %dx%d smaller than a tile", scene_h, scene_w);
    const long long n_patches = (long long)n_tiles * t0 * t1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ws) {
        case 3:  launch_descriptor_fast<3>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        case 5:  launch_descriptor_fast<5>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        case 7:  launch_descriptor_fast<7>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        case 9:  launch_descriptor_fast<9>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        case 11: launch_descriptor_fast<11>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        case 13: launch_descriptor_fast<13>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        case 15: launch_descriptor_fast<15>(scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, desc_bf16_dev, stat_dev, st); break;
        default: {
            const int warps = 8;
            dm_descriptor_kernel<<<dm_div_up(n_patches, warps), warps * 32, 0, st>>>(
                scene_dev, pitch, origin_yx_dev, n_patches, t0, t1, ws, dm_kpad(ws),
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
