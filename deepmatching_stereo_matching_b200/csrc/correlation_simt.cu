// CUDA-core correlation: the exact reference engine for the tcgen05 kernel and the path
// for patch grids the tensor-core tiling does not cover (P not a multiple of 128).
// Replaces the P calls of cv2.matchTemplate in
// Correlation_map._create_simple_initial_co_map (misc/Correlation_map.py:69-87).
//
// raw[t][p][q] = ZNCC( desc1[t][p][:], desc2[t][q][:] ); bf16 operands hold exact
// integers, the fp32 accumulation is exact (|sum| <= 225*255^2 < 2^24 for ws <= 15).  The tensor-core
// kernel gets the -S1' S2'/K term from three correction entries of the descriptor rows and may
// differ from this engine by the rounding of their accumulation (a few ulp of the accumulator).
#include "dm_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 32;

__global__ void __launch_bounds__(256)
dm_correlation_simt_kernel(const __nv_bfloat16* __restrict__ d1, const dm_stat* __restrict__ st1,
                           const __nv_bfloat16* __restrict__ d2, const dm_stat* __restrict__ st2,
                           int P, int kpad, int normed, int slot0, int slot1, int slot2, float* __restrict__ raw) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tile = blockIdx.z;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const __nv_bfloat16* a = d1 + (size_t)tile * P * kpad;
    const __nv_bfloat16* b = d2 + (size_t)tile * P * kpad;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < kpad; k0 += TK) {
        // 64 rows x 32 k per operand = 2048 values, 8 per thread; k fastest for coalescing
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int e = it * 256 + threadIdx.x;
            int r = e >> 5, k = e & 31;
            float va = 0.f, vb = 0.f;
            // the correction slots (descriptors.cu) are for the tensor-core kernel: this engine keeps the dot
            // product an exact integer and applies -S1' S2'/K with one FMA in the epilogue
            const int kk = k0 + k;
            const bool special = kk == slot0 || kk == slot1 || kk == slot2;
            if (m0 + r < P && !special) va = __bfloat162float(a[(size_t)(m0 + r) * kpad + kk]);
            if (n0 + r < P && !special) vb = __bfloat162float(b[(size_t)(n0 + r) * kpad + kk]);
            As[k][r] = va;
            Bs[k][r] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { av[u] = As[k][ty * 4 + u]; bv[u] = Bs[k][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        __syncthreads();
    }
    const dm_stat* s1 = st1 + (size_t)tile * P;
    const dm_stat* s2 = st2 + (size_t)tile * P;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int p = m0 + ty * 4 + u;
        if (p >= P) continue;
        dm_stat sp = s1[p];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            int q = n0 + tx * 4 + v;
            if (q >= P) continue;
            dm_stat sq = s2[q];
            float z = dm_zncc_partial(acc[u][v], sp.x, sq.z, normed ? sq.y : 1.0f);
            raw[((size_t)tile * P + p) * P + q] = dm_zncc_finish(z, sp.y, sp.y == 0.f, normed != 0);
        }
    }
}

}  // namespace

void dm_desc_special_slots(int ws, int k[3]);

int dm_correlation_simt(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                        int n_tiles, int p, int kpad, int ws, int method, float* raw, cudaStream_t stream) {
    int slot[3];
    dm_desc_special_slots(ws, slot);
    dim3 grid(dm_div_up(p, TN), dm_div_up(p, TM), n_tiles);
    DM_REQUIRE(n_tiles <= 65535, DM_ERR_INVALID, "dm_correlation: more than 65535 tiles per call");
    dm_correlation_simt_kernel<<<grid, 256, 0, stream>>>(
        (const __nv_bfloat16*)desc1, (const dm_stat*)stat1, (const __nv_bfloat16*)desc2, (const dm_stat*)stat2,
        p, kpad, method == DM_TM_CCOEFF_NORMED, slot[0], slot[1], slot[2], raw);
    DM_LAUNCH_CHECK();
    return DM_OK;
}
