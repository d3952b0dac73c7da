// Bilateral filter of an 8-bit plane: the live branch of the reference's post-process,
// optimize_looper.py:76-77 -- cv2.bilateralFilter(d_map.astype('uint8'), 2*exclusion+1, sigma, sigma).
// SURVEY.md section 8(f) row 3.
//
// Restates OpenCV's own 8-bit single-channel algorithm (modules/imgproc/src/bilateral_filter
// .dispatch.cpp / .simd.hpp, the non-IPP path; checked bit for bit against cv2 4.13 with IPP
// switched off, see tests/golden/make_golden_bilateral.py):
//   radius = d / 2 (d <= 0: round(1.5 * sigma_space)), at least 1; sigma <= 0 counts as 1;
//   colour weight  cw[i] = (float)exp(i*i * (float)(-0.5 / sigma_color^2)),  i = |v - v0|
//   space weight   sw[k] = (float)exp(r*r * (float)(-0.5 / sigma_space^2))   for r = hypot(i, j) <= radius,
//                  k running over i = -radius..radius (rows), j = -radius..radius (columns)
//   border         BORDER_REFLECT_101
//   out            = cvRound(sum_k v_k * w_k / sum_k w_k), w_k = sw[k] * cw[|v_k - v0|],
//                  float32, accumulated in k order, products and sums rounded separately.
// Builds of OpenCV that route small windows through IPP differ from this by one grey level
// on part of the pixels; that is OpenCV against itself, not this kernel.
#include <math.h>

#include <vector>

#include "dm_common.cuh"

namespace {

constexpr int BIL_MAX_RADIUS = 15;
constexpr int BIL_MAX_K = (2 * BIL_MAX_RADIUS + 1) * (2 * BIL_MAX_RADIUS + 1);

struct BilateralTables {
    float cw[256];
    float sw[BIL_MAX_K];
    int8_t di[BIL_MAX_K], dj[BIL_MAX_K];
    int maxk;
};

__device__ __forceinline__ int reflect101(int p, int n) {
    // BORDER_REFLECT_101 for |overshoot| < n (radius < n is required by the entry point)
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256)
dm_bilateral_u8_kernel(const uint8_t* __restrict__ src, int H, int W, const BilateralTables* __restrict__ tab, uint8_t* __restrict__ dst) {
    __shared__ float cw[256];
    __shared__ float sw[BIL_MAX_K];
    __shared__ int8_t di[BIL_MAX_K], dj[BIL_MAX_K];
    const int maxk = tab->maxk;
    for (int t = threadIdx.x; t < 256; t += blockDim.x) cw[t] = tab->cw[t];
    for (int t = threadIdx.x; t < maxk; t += blockDim.x) { sw[t] = tab->sw[t]; di[t] = tab->di[t]; dj[t] = tab->dj[t]; }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int v0 = src[(size_t)y * W + x];
    float sum = 0.0f, wsum = 0.0f;
    for (int k = 0; k < maxk; ++k) {
        const int yy = reflect101(y + di[k], H), xx = reflect101(x + dj[k], W);
        const int v = src[(size_t)yy * W + xx];
        const float w = __fmul_rn(sw[k], cw[abs(v - v0)]);
        sum = __fadd_rn(sum, __fmul_rn((float)v, w));
        wsum = __fadd_rn(wsum, w);
    }
    dst[(size_t)y * W + x] = (uint8_t)__float2int_rn(__fdiv_rn(sum, wsum));      // cvRound: half to even
}

}  // namespace

extern "C" int dm_bilateral_u8(const uint8_t* src_dev, int h, int w, int d, double sigma_color, double sigma_space,
                               uint8_t* dst_dev, void* stream) {
    DM_REQUIRE(src_dev && dst_dev && src_dev != dst_dev, DM_ERR_INVALID, "dm_bilateral_u8: null or aliased planes");
    DM_REQUIRE(h > 0 && w > 0, DM_ERR_INVALID, "dm_bilateral_u8: empty plane");
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = d <= 0 ? (int)lrint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    DM_REQUIRE(radius <= BIL_MAX_RADIUS, DM_ERR_UNSUPPORTED, "dm_bilateral_u8: radius %d exceeds %d", radius, BIL_MAX_RADIUS);
    DM_REQUIRE(radius < h && radius < w, DM_ERR_UNSUPPORTED, "dm_bilateral_u8: plane %dx%d smaller than the window radius %d", h, w, radius);
    const float gc = (float)(-0.5 / (sigma_color * sigma_color)), gs = (float)(-0.5 / (sigma_space * sigma_space));
    BilateralTables t;
    for (int i = 0; i < 256; ++i) t.cw[i] = (float)exp((double)((float)(i * i) * gc));
    t.maxk = 0;
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            t.sw[t.maxk] = (float)exp(r * r * (double)gs);
            t.di[t.maxk] = (int8_t)i; t.dj[t.maxk] = (int8_t)j;
            ++t.maxk;
        }
    cudaStream_t st = (cudaStream_t)stream;
    BilateralTables* dev = nullptr;
    DM_CUDA_CHECK(cudaMallocAsync((void**)&dev, sizeof(BilateralTables), st));
    // the table is small and the copy is stream ordered; the pageable source is read before the call returns
    DM_CUDA_CHECK(cudaMemcpyAsync(dev, &t, sizeof(BilateralTables), cudaMemcpyHostToDevice, st));
    DM_CUDA_CHECK(cudaStreamSynchronize(st));
    dim3 grid(dm_div_up(w, 32), dm_div_up(h, 8));
    dm_bilateral_u8_kernel<<<grid, 256, 0, st>>>(src_dev, h, w, dev, dst_dev);
    DM_LAUNCH_CHECK();
    DM_CUDA_CHECK(cudaFreeAsync(dev, st));
    return DM_OK;
}
