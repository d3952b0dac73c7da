// Bottom-up correlation pyramid: per-slice min-max + rectification, and the level
// transition (3x3/s2/p1 max-pool, average of the four children, **1.4).
// Replaces Feature_value.min_max (misc/Feature_value.py:32-37),
// Correlation_map._rectification (misc/Correlation_map.py:158-159) and
// Correlation_map._aggregation (misc/Correlation_map.py:89-130).
//
// Both kernels are HBM-bound: every level-k value is read once and every level-(k+1)
// value written once.  The aggregation stages the four child slices of a parent in shared
// memory with bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier),
// reads them back with 128-bit loads and hands the one-column halo between lanes with a
// warp shuffle.
#include <atomic>
#include <cstdlib>

#include "dm_common.cuh"

namespace {

// ------------------------------------------------------------------ min-max + rectify
template <int CACHE>   // CACHE float4 per thread held in registers (row read once from HBM)
__global__ void __launch_bounds__(256)
dm_minmax_rectify_kernel(const float* raw, int q, float* norm,
                         float* rect, float* __restrict__ rowmin, float* __restrict__ rowmax) {
    __shared__ float red[2][8];
    const size_t row = blockIdx.x;
    const float* src = raw + row * (size_t)q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    float4 buf[CACHE > 0 ? CACHE : 1];
    if (CACHE > 0) {
        const int q4 = q >> 2;
#pragma unroll
        for (int c = 0; c < CACHE; ++c) {
            int i = c * 256 + threadIdx.x;
            if (i < q4) {
                buf[c] = reinterpret_cast<const float4*>(src)[i];   // plain load: norm/rect may alias raw
                mn = dm_min_nan(dm_min_nan(mn, buf[c].x), dm_min_nan(buf[c].y, dm_min_nan(buf[c].z, buf[c].w)));
                mx = dm_max_nan(dm_max_nan(mx, buf[c].x), dm_max_nan(buf[c].y, dm_max_nan(buf[c].z, buf[c].w)));
            }
        }
    } else {
        for (int i = threadIdx.x; i < q; i += 256) {
            float v = src[i];
            mn = dm_min_nan(mn, v);
            mx = dm_max_nan(mx, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = dm_min_nan(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = dm_max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red[0][warp] = mn; red[1][warp] = mx; }
    __syncthreads();
    mn = red[0][0]; mx = red[1][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { mn = dm_min_nan(mn, red[0][w]); mx = dm_max_nan(mx, red[1][w]); }
    if (threadIdx.x == 0) {
        if (rowmin) rowmin[row] = mn;
        if (rowmax) rowmax[row] = mx;
    }
    const float rinv = dm_range_inv(mn, mx);
    float* dn = norm ? norm + row * (size_t)q : nullptr;
    float* dr = rect ? rect + row * (size_t)q : nullptr;
    if (CACHE > 0) {
        const int q4 = q >> 2;
#pragma unroll
        for (int c = 0; c < CACHE; ++c) {
            int i = c * 256 + threadIdx.x;
            if (i < q4) {
                float4 v = buf[c];
                v.x = dm_normalize(v.x, mn, mx, rinv); v.y = dm_normalize(v.y, mn, mx, rinv);
                v.z = dm_normalize(v.z, mn, mx, rinv); v.w = dm_normalize(v.w, mn, mx, rinv);
                if (dn) reinterpret_cast<float4*>(dn)[i] = v;
                if (dr) {
                    v.x = dm_rectify(v.x); v.y = dm_rectify(v.y); v.z = dm_rectify(v.z); v.w = dm_rectify(v.w);
                    reinterpret_cast<float4*>(dr)[i] = v;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < q; i += 256) {
            float v = dm_normalize(src[i], mn, mx, rinv);      // plain load: src may alias norm
            if (dn) dn[i] = v;
            if (dr) dr[i] = dm_rectify(v);
        }
    }
}

// ------------------------------------------------------------------ aggregation
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// in [n][A][B][C][D] -> out [n][A/2][B/2][C/2][D/2].  One CTA owns `pp` consecutive parents
// (flattened over n, A/2, B/2) and a band of `rb` output rows of their maps.
// Requirements of this kernel: D % 4 == 0 (16-byte rows), C even.
__global__ void __launch_bounds__(256)
dm_aggregate_kernel(const float* __restrict__ in, long long n_parents, int A, int B, int C, int D,
                    int pp, int rb, int rect, int merged, dm_fastdiv fd_xp, dm_fastdiv fd_oc, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    dm_pdl_wait();                  // the level below comes from the kernel in front
    dm_pdl_launch_dependents();
    const int hA = A >> 1, hB = B >> 1, oc = C >> 1, od = D >> 1;
    const long long parent0 = (long long)blockIdx.x * pp;
    const int npar = (int)min((long long)pp, n_parents - parent0);
    const int y0 = blockIdx.y * rb;                       // first output row of the band
    const int y1 = min(oc, y0 + rb);
    const int r0 = max(0, 2 * y0 - 1);                    // first input row staged
    const int nrows = 2 * y1 - r0;                        // input rows r0 .. 2*y1-1
    const size_t slice = (size_t)C * D;
    const uint32_t chunk_bytes = (uint32_t)(nrows * D * sizeof(float));
    const int chunk = nrows * D;                          // floats per staged child band

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && merged) {
        // whole slices, all parents of the CTA in one parent row I: the children (2I + c, 2J0 .. 2J0 + 2 pp - 1)
        // of one child row are 2 pp consecutive slices of the input -- TWO bulk copies per CTA instead of 4 pp,
        // and one index decode instead of pp (a 512^2 / image_size 32 level 1: 2 x 16 KiB instead of 32 x 1 KiB,
        // each behind four 64-bit divisions of its own)
        const int J0 = (int)(parent0 % hB);
        const long long t = parent0 / hB;
        const int I = (int)(t % hA);
        const long long nn = t / hA;
        const uint32_t run_bytes = chunk_bytes * 2u * (uint32_t)npar;
        mbar_expect_tx(&bar, 2u * run_bytes);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const size_t child = ((size_t)nn * A + (2 * I + c)) * B + 2 * J0;
            bulk_g2s(smem + (size_t)c * 2 * npar * chunk, in + child * slice, run_bytes, &bar);
        }
    } else if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, chunk_bytes * 4u * (uint32_t)npar);
        for (int k = 0; k < npar; ++k) {
            long long par = parent0 + k;
            int J = (int)(par % hB);
            long long t = par / hB;
            int I = (int)(t % hA);
            long long nn = t / hA;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                size_t child = ((size_t)nn * A + (2 * I + (ch >> 1))) * B + (2 * J + (ch & 1));
                bulk_g2s(smem + (size_t)(k * 4 + ch) * chunk, in + child * slice + (size_t)r0 * D, chunk_bytes, &bar);
            }
        }
    }
    mbar_wait(&bar, 0);

    const int xpairs = D >> 2;                            // two outputs per item
    const int rows = y1 - y0;
    const int items = npar * rows * xpairs;
    const int lane = threadIdx.x & 31;
    const float NEG = -CUDART_INF_F;
    for (int it0 = 0; it0 < items; it0 += 256) {
        const int it = it0 + threadIdx.x;
        const bool active = it < items;
        int xp = 0, y = 0, k = 0;
        if (active) {
            // item -> (parent k, row y, pair of columns xp) by exact multiply-shift divisions (four
            // runtime-divisor / and % per item were 40 % of the kernel's instructions).  Several
            // parents per CTA only occur with one band, i.e. rows == C / 2.
            const uint32_t t = dm_fd_div((uint32_t)it, fd_xp);
            xp = it - (int)t * xpairs;
            k = pp > 1 ? (int)dm_fd_div(t, fd_oc) : 0;
            y = (int)t - k * rows;
        }
        const int yy = y0 + y;                            // output row
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            // staged slice of child ch of parent k: parent-major, or (merged) the two child rows one after the other
            const int slot = merged ? (ch >> 1) * 2 * npar + 2 * k + (ch & 1) : k * 4 + ch;
            const float* base = smem + (size_t)slot * chunk;
            float m0 = NEG, m1 = NEG;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int r = 2 * yy + dy;                // input row; r < C always, r >= 0 unless yy == 0
                const bool rowok = active && r >= 0;
                float4 v = make_float4(NEG, NEG, NEG, NEG);
                if (rowok) v = *reinterpret_cast<const float4*>(base + (size_t)(r - r0) * D + 4 * xp);
                float left = __shfl_up_sync(0xffffffffu, v.w, 1);
                if (xp == 0) left = NEG;
                else if (lane == 0 && rowok) left = base[(size_t)(r - r0) * D + 4 * xp - 1];
                m0 = dm_max_nan(m0, dm_max_nan(dm_max_nan(left, v.x), v.y));
                m1 = dm_max_nan(m1, dm_max_nan(dm_max_nan(v.y, v.z), v.w));
            }
            // add order ((ul + ur) + ll) + lr  (misc/Correlation_map.py:122)
            sum0 = (ch == 0) ? m0 : __fadd_rn(sum0, m0);
            sum1 = (ch == 0) ? m1 : __fadd_rn(sum1, m1);
        }
        if (active) {
            float2 o;
            o.x = __fmul_rn(sum0, 0.25f);
            o.y = __fmul_rn(sum1, 0.25f);
            if (rect) { o.x = dm_rectify(o.x); o.y = dm_rectify(o.y); }
            size_t off = ((size_t)(parent0 + k) * oc + yy) * od + 2 * xp;
            *reinterpret_cast<float2*>(out + off) = o;
        }
    }
}

// Fallback for tiny or odd shapes (top of the pyramid: C or D == 2, D % 4 != 0).
__global__ void __launch_bounds__(256)
dm_aggregate_small_kernel(const float* __restrict__ in, long long n_out, int A, int B, int C, int D,
                          int rect, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_out) return;
    const int hA = A >> 1, hB = B >> 1, oc = C >> 1, od = D >> 1;
    int x = (int)(idx % od);
    long long t = idx / od;
    int y = (int)(t % oc); t /= oc;
    int J = (int)(t % hB); t /= hB;
    int I = (int)(t % hA);
    long long nn = t / hA;
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const float* src = in + (((size_t)nn * A + (2 * I + (ch >> 1))) * B + (2 * J + (ch & 1))) * (size_t)C * D;
        float m = -CUDART_INF_F;
        for (int dy = -1; dy <= 1; ++dy) {
            int r = 2 * y + dy;
            if (r < 0 || r >= C) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                int c = 2 * x + dx;
                if (c < 0 || c >= D) continue;
                m = dm_max_nan(m, __ldg(src + (size_t)r * D + c));
            }
        }
        sum = (ch == 0) ? m : __fadd_rn(sum, m);
    }
    const float avg = __fmul_rn(sum, 0.25f);
    out[idx] = rect ? dm_rectify(avg) : avg;
}

}  // namespace

int dm_minmax_rectify_impl(const float* raw, long long rows, int q, float* norm, float* rect,
                           float* rowmin, float* rowmax, cudaStream_t stream) {
    DM_REQUIRE(rows > 0 && q > 0 && rows < (1LL << 31), DM_ERR_INVALID, "dm_minmax_rectify: bad shape");
    const bool vec = (q % 4 == 0) && ((((uintptr_t)raw | (uintptr_t)norm | (uintptr_t)rect) & 15) == 0);
    if (vec && q <= 4 * 256 * 1)
        dm_minmax_rectify_kernel<1><<<(unsigned)rows, 256, 0, stream>>>(raw, q, norm, rect, rowmin, rowmax);
    else if (vec && q <= 4 * 256 * 4)
        dm_minmax_rectify_kernel<4><<<(unsigned)rows, 256, 0, stream>>>(raw, q, norm, rect, rowmin, rowmax);
    else if (vec && q <= 4 * 256 * 16)
        dm_minmax_rectify_kernel<16><<<(unsigned)rows, 256, 0, stream>>>(raw, q, norm, rect, rowmin, rowmax);
    else
        dm_minmax_rectify_kernel<0><<<(unsigned)rows, 256, 0, stream>>>(raw, q, norm, rect, rowmin, rowmax);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_minmax_rectify(const float* raw_dev, long long rows, int q, float* norm_dev, float* rect_dev,
                                 float* rowmin_dev, float* rowmax_dev, void* stream) {
    return dm_minmax_rectify_impl(raw_dev, rows, q, norm_dev, rect_dev, rowmin_dev, rowmax_dev, (cudaStream_t)stream);
}

extern "C" int dm_aggregate(const float* in_dev, int n, int a, int b, int c, int d, int rectify, float* out_dev, void* stream) {
    DM_REQUIRE(n > 0 && a >= 2 && b >= 2 && c >= 2 && d >= 2 && !(a & 1) && !(b & 1) && !(c & 1) && !(d & 1),
               DM_ERR_INVALID, "dm_aggregate: level shape (%d,%d,%d,%d) must be even and >= 2", a, b, c, d);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_parents = (long long)n * (a / 2) * (b / 2);
    const int oc = c / 2;
    const bool fast = (d % 4 == 0) && ((((uintptr_t)in_dev) & 15) == 0) && ((((uintptr_t)out_dev) & 7) == 0);
    if (!fast) {
        const long long n_out = n_parents * oc * (d / 2);
        dm_aggregate_small_kernel<<<dm_div_up(n_out, 256), 256, 0, st>>>(in_dev, n_out, a, b, c, d, rectify, out_dev);
        DM_LAUNCH_CHECK();
        return DM_OK;
    }
    // shared-memory plan: <= 64 KiB per CTA so three CTAs stay resident per SM
    // (32 KiB for the small upper levels: a CTA there lives only a few microseconds, and six resident
    //  CTAs hide the load -> compute -> exit chain better than three: 0.234 -> 0.205 ms for level 1 -> 2 of C2)
    const size_t full = (size_t)4 * c * d * sizeof(float);          // four whole child slices
    static const int budget_kb = getenv("DM_AGG_BUDGET_KB") ? atoi(getenv("DM_AGG_BUDGET_KB")) : 0;     // measurement aid
    const size_t budget = budget_kb > 0 && full <= 16 * 1024 ? (size_t)budget_kb * 1024 : (full <= 16 * 1024 ? 32 * 1024 : 64 * 1024);
    int pp = 1, rb = oc;
    if (full <= budget) {
        pp = (int)(budget / full);
        if (pp > 64) pp = 64;
        // keep at least ~4 CTAs per SM worth of blocks when the level is small
        while (pp > 1 && n_parents / pp < 148 * 4) pp >>= 1;
    } else {
        // band of rb output rows needs 2*rb+1 input rows of each child
        rb = (int)((budget / (4 * (size_t)d * sizeof(float)) - 1) / 2);
        if (rb < 1) rb = 1;
        if (rb > oc) rb = oc;
    }
    const int bands = dm_div_up(oc, rb);
    // one band and the parents of a CTA inside one parent row: the staged slices are two contiguous runs of the input
    static const bool no_merge = getenv("DM_AGG_NO_MERGE") != nullptr;     // measurement aid: one bulk copy per child slice
    int merged = 0;
    if (bands == 1 && !no_merge) {
        while (pp > 1 && (b / 2) % pp != 0) --pp;
        merged = 1;
    }
    const int rows_staged = (bands == 1) ? c : 2 * rb + 1;
    const size_t smem = (size_t)pp * 4 * rows_staged * d * sizeof(float);
    DM_REQUIRE(smem <= 200 * 1024, DM_ERR_UNSUPPORTED, "dm_aggregate: row of %d floats too wide for shared memory", d);
    {   // the attribute belongs to the device: one flag per device (several devices per process, dm_multi_*)
        static std::atomic<bool> attr_set[64];
        int dev = 0;
        DM_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
            DM_CUDA_CHECK(cudaFuncSetAttribute(dm_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
        }
    }
    dim3 grid((unsigned)dm_div_up(n_parents, pp), bands);
    dm_launch_dep(DM_PDL_UPPER, dm_aggregate_kernel, grid, dim3(256), smem, st, in_dev, n_parents, a, b, c, d, pp, rb, rectify, merged,
                  dm_make_fastdiv((uint32_t)(d >> 2)), dm_make_fastdiv((uint32_t)oc), out_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}
