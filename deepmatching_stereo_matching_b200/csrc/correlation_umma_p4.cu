// tcgen05 correlation engine, pooled epilogue, SIXTEEN epilogue warps (sm_100a).
//
// Same contraction, same pooled ZNCC epilogue and same outputs as MODE_POOL of
// correlation_umma.cu: the P calls of cv2.matchTemplate in
// Correlation_map._create_simple_initial_co_map (misc/Correlation_map.py:69-87) fused with
// the 3x3/s2/p1 max-pool of Correlation_map._aggregation (misc/Correlation_map.py:100-103)
// and the per-patch min / max of Feature_value.min_max (misc/Feature_value.py:32-37).
//
// Why: with 8 epilogue warps (2 per scheduler) the epilogue was latency-bound -- ncu showed
// 34 % issue-slot use, 32 % of the warp time on the TMEM-load scoreboard and 25 % on the
// shared-memory one, and ptxas sinks the software-pipelined tcgen05.ld towards its use so
// the prefetch distance cannot be forced from the source.  Four warps per scheduler hide
// those latencies by switching warps instead.
//
// Work item = 256 patches of one tile (two M=128 accumulator halves) x all P positions,
// swept in N-tiles of 128 positions.  Persistent grid, one CTA per SM, 20 warps in five
// warpgroups (setmaxnreg: 40 registers for warpgroup 0, 104 for the epilogue):
//   warp 0       TMA producer  A (2 halves x KB boxes, once per item), ring of B boxes,
//                              ring of 1 KiB column tables
//   warp 1       MMA issuer    tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16, operands
//                              from 128-byte-swizzled shared memory; TMEM: 2 halves x 2
//                              accumulator stages x 128 columns.  Padding-only K steps are
//                              skipped (ws = 15: 15 of 16).
//   warps 4..19  epilogue      warp w may touch TMEM lanes 32*(w % 4)..+31, so four warps share
//                              a lane quarter and split every map row's columns four ways
//                              (q4 = (w - 4) / 4).  A thread owns two patch rows (lane l of
//                              both halves) and D/4 columns of every map row; the 3x3 pooling,
//                              the running row minimum and the halo row carried between
//                              N-tiles are thread-local; the one halo column left of the
//                              thread's range comes from a 1-column TMEM load.  Steps of 8
//                              columns, the next step's TMEM loads and column parameters in
//                              flight during the math.  Every 8 pooled values per row leave
//                              through an XOR-swizzled shared-memory transpose so that a store
//                              instruction writes 16 rows x 32 B (whole sectors).
#include "dm_common.cuh"
#include "dm_internal.h"
#include "umma.cuh"

namespace {

constexpr int BM = 128;                 // rows per accumulator half = TMEM lanes
constexpr int HALVES = 2;
constexpr int BN = 128;                 // positions per N-tile
constexpr int BK = 64;                  // bf16 per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int MAX_KB = 4;               // kpad <= 256
constexpr int BOX_BYTES = BM * BK * 2;  // 16 KiB: one [128 rows x 64] bf16 box
constexpr int EPI_WARPS = 16;
constexpr int NSPLIT = 4;               // column split of a map row between the warps of a lane quarter
constexpr int FIRST_EPI_WARP = 4;       // warps 0..3 = warpgroup 0: TMA producer, MMA issuer, 2 idle
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int SW = 8;                   // columns per epilogue step
constexpr int NSTEP = (BN / NSPLIT) / SW;   // 4
constexpr int STG_FLOATS = 32 * 8;      // one staging region: 32 rows x 8 floats, float4 slots XOR-swizzled
constexpr int STG_BYTES = 2 * STG_FLOATS * 4;   // per epilogue warp: one region per accumulator half
constexpr int CS_BYTES = (BN / 2) * 16; // column table of one N-tile: 64 x {sk0, sk1, inv0, inv1}
constexpr int CS_STAGES = 4;

constexpr size_t OFF_A = 0;
constexpr size_t OFF_B = OFF_A + (size_t)HALVES * MAX_KB * BOX_BYTES;
constexpr size_t OFF_STG = OFF_B + (size_t)STAGES * BOX_BYTES;
constexpr size_t OFF_CS = OFF_STG + (size_t)EPI_WARPS * STG_BYTES;
constexpr size_t OFF_BAR = OFF_CS + CS_STAGES * CS_BYTES;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + OFF_BAR + 256;

struct Params {
    const dm_stat* stat1;       // [n*P] float4 {S', inv, S'/K, mean}
    const float4* cstat2;       // [n*P/2] {S'/K even, S'/K odd, inv even, inv odd} of image 2
    int n_items, P, KB, ksteps, items_per_tile;
    float* pooled;              // [n][P][P/4]
    float* rowmin; float* rowmax;   // [n][P][4] partial min / max of the four column quarters
};

enum { MODE_POOL = 1, MODE_NULL = 2 };   // MODE_NULL: drain TMEM only (measurement aid)

template <int MODE, int D, bool NORMED>      // D = positions per map row (T1): 32, 64 or 128
__global__ void __launch_bounds__(THREADS, 1)
dm_correlation_p4_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smemA = smem + OFF_A;
    uint8_t* smemB = smem + OFF_B;
    float* smemStg = reinterpret_cast<float*>(smem + OFF_STG);
    uint8_t* smemCs = smem + OFF_CS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;
    uint64_t* b_empty = b_full + STAGES;
    uint64_t* t_full = b_empty + STAGES;
    uint64_t* t_empty = t_full + 2;
    uint64_t* c_full = t_empty + 2;
    uint64_t* c_empty = c_full + CS_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + CS_STAGES);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int P = prm.P, KB = prm.KB, NT = P / BN;

    if (threadIdx.x == 0) {
        umma::mbar_init(a_full, 1);
        umma::mbar_init(a_empty, 1);
        for (int s = 0; s < STAGES; ++s) { umma::mbar_init(b_full + s, 1); umma::mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { umma::mbar_init(t_full + s, 1); umma::mbar_init(t_empty + s, EPI_WARPS); }
        for (int s = 0; s < CS_STAGES; ++s) { umma::mbar_init(c_full + s, 1); umma::mbar_init(c_empty + s, EPI_WARPS); }
        umma::fence_barrier_init();
        umma::tma_prefetch_desc(&mapA);
        umma::tma_prefetch_desc(&mapB);
    }
    if (warp == 1) {
        umma::tmem_alloc(tmem_slot, TMEM_COLS);
        umma::tmem_relinquish();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp < FIRST_EPI_WARP) {
        umma::reg_dealloc<40>();
        if (warp == 0) {
            // -------------------------------------------------------- TMA producer (converged warp, elected issue)
            int bs = 0; uint32_t bph = 0, aph = 0; int cst = 0; uint32_t cph = 0;
            for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
                const int tile = item / prm.items_per_tile;
                const int row0 = tile * P + (item - tile * prm.items_per_tile) * (HALVES * BM);
                umma::mbar_wait(a_empty, aph ^ 1);
                if (umma::elect_one()) {
                    umma::mbar_expect_tx(a_full, (uint32_t)(HALVES * KB * BOX_BYTES));
                    for (int h = 0; h < HALVES; ++h)
                        for (int kb = 0; kb < KB; ++kb)
                            umma::tma_load_2d(smemA + (size_t)(h * MAX_KB + kb) * BOX_BYTES, &mapA, kb * BK, row0 + h * BM, a_full);
                }
                __syncwarp();
                aph ^= 1;
                for (int j = 0; j < NT; ++j) {
                    for (int kb = 0; kb < KB; ++kb) {
                        umma::mbar_wait(b_empty + bs, bph ^ 1);
                        if (umma::elect_one()) {
                            umma::mbar_expect_tx(b_full + bs, (uint32_t)BOX_BYTES);
                            umma::tma_load_2d(smemB + (size_t)bs * BOX_BYTES, &mapB, kb * BK, tile * P + j * BN, b_full + bs);
                        }
                        __syncwarp();
                        if (++bs == STAGES) { bs = 0; bph ^= 1; }
                    }
                    umma::mbar_wait(c_empty + cst, cph ^ 1);
                    if (umma::elect_one()) {
                        umma::mbar_expect_tx(c_full + cst, (uint32_t)CS_BYTES);
                        umma::bulk_g2s(smemCs + (size_t)cst * CS_BYTES, prm.cstat2 + ((size_t)tile * P + (size_t)j * BN) / 2, CS_BYTES, c_full + cst);
                    }
                    __syncwarp();
                    if (++cst == CS_STAGES) { cst = 0; cph ^= 1; }
                }
            }
        } else if (warp == 1) {
            // -------------------------------------------------------- MMA issuer (converged warp, elected issue)
            constexpr uint32_t idesc = umma::instr_desc_bf16(BM, BN);
            constexpr int KPB = BK / UMMA_K;
            int bs = 0; uint32_t bph = 0, aph = 0; int acc = 0; uint32_t accph = 0;
            for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
                umma::mbar_wait(a_full, aph);
                aph ^= 1;
                for (int j = 0; j < NT; ++j) {
                    umma::mbar_wait(t_empty + acc, accph ^ 1);
                    umma::tc_fence_after();
                    const uint32_t d0 = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < KB; ++kb) {
                        umma::mbar_wait(b_full + bs, bph);
                        umma::tc_fence_after();
                        const uint64_t bdesc = umma::smem_desc_sw128(smemB + (size_t)bs * BOX_BYTES);
                        const uint64_t adesc0 = umma::smem_desc_sw128(smemA + (size_t)kb * BOX_BYTES);
                        const int nk = prm.ksteps - kb * KPB;       // padding-only K steps are skipped
#pragma unroll
                        for (int k = 0; k < KPB; ++k) {             // +32 B per K step inside the swizzle atom
                            if (k < nk) {
#pragma unroll
                                for (int h = 0; h < HALVES; ++h)
                                    if (umma::elect_one())
                                        umma::mma_bf16(d0 + (uint32_t)(h * 2 * BN), adesc0 + (uint64_t)(h * ((MAX_KB * BOX_BYTES) >> 4)) + 2 * k,
                                                       bdesc + 2 * k, idesc, (kb | k) != 0);
                            }
                        }
                        if (umma::elect_one()) umma::mma_commit(b_empty + bs);
                        if (++bs == STAGES) { bs = 0; bph ^= 1; }
                    }
                    if (umma::elect_one()) umma::mma_commit(t_full + acc);
                    if (++acc == 2) { acc = 0; accph ^= 1; }
                }
                if (umma::elect_one()) umma::mma_commit(a_empty);
            }
        }
    } else {
        umma::reg_alloc<104>();     // 640 threads start at 96; warpgroup 0 frees 128 x 56, the four epilogue groups take 512 x 8
        // ------------------------------------------------------------ epilogue
        const int e = warp - FIRST_EPI_WARP;
        const int quarter = warp & 3;                   // TMEM lanes a warp may touch: 32*(warp_id % 4)..+31
        const int q4 = e >> 2;                          // which quarter of every map row's columns
        constexpr int DQ = D / NSPLIT;                  // columns of a map row handled by this thread (8, 16, 32)
        constexpr int HWQ = DQ / 2;                     // pooled outputs per map row per thread
        constexpr int SPR = DQ / SW;                    // steps per map row (1, 2, 4)
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        float* stgA = smemStg + (size_t)e * (2 * STG_FLOATS);
        float* stgB = stgA + STG_FLOATS;
        const int wsw = (lane >> 2) & 1;                // XOR swizzle of this lane's own staging row
        int acc = 0; uint32_t accph = 0; int cst = 0; uint32_t cph = 0;
        for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
            const int tile = item / prm.items_per_tile;
            const size_t wrow = (size_t)tile * P + (size_t)(item - tile * prm.items_per_tile) * (HALVES * BM) + quarter * 32;
            const size_t prowA = wrow + lane, prowB = prowA + BM;
            const dm_stat s1A = prm.stat1[prowA], s1B = prm.stat1[prowB];
            const float ns1A = -s1A.x, ns1B = -s1B.x;
            // st[] = horizontally pooled previous map row (odd rows) / running vertical max (even rows)
            float stA[HWQ], stB[HWQ];
            float rminA = CUDART_INF_F, rmaxA = -CUDART_INF_F, rminB = CUDART_INF_F, rmaxB = -CUDART_INF_F;
#pragma unroll
            for (int i = 0; i < HWQ; ++i) { stA[i] = -CUDART_INF_F; stB[i] = -CUDART_INF_F; }
            const size_t ostride = (size_t)(P / 4);
            float* woutA = prm.pooled + wrow * ostride;
            float* woutB = woutA + (size_t)BM * ostride;
            // 8 staged floats per lane -> the warp writes 16 rows x 32 B per store instruction.
            // The 8 floats are segments of `seg` floats, `segstride` floats apart in the output.
            auto flush8 = [&](const float* stg, float* wout, size_t col, int seg, size_t segstride) {
                __syncwarp();
                const int f = lane & 1, f4 = f * 4, row = lane >> 1;
                const uint32_t sa = umma::smem_u32(stg + row * 8 + ((f ^ ((row >> 2) & 1)) << 2));
                const float4 v0 = umma::lds128(sa);                                 // rows 0..15
                const float4 v1 = umma::lds128(sa + 16 * 8 * 4);                    // rows 16..31 (same swizzle bit)
                float* dst = wout + (size_t)row * ostride + col + (size_t)(f4 / seg) * segstride + (f4 % seg);
                *reinterpret_cast<float4*>(dst) = v0;
                *reinterpret_cast<float4*>(dst + (size_t)16 * ostride) = v1;
                __syncwarp();
            };
            for (int j = 0; j < NT; ++j) {
                umma::mbar_wait(c_full + cst, cph);
                umma::mbar_wait(t_full + acc, accph);
                umma::tc_fence_after();
                const uint32_t tA = tmem_base + lane_addr + (uint32_t)((0 * 2 + acc) * BN);
                const uint32_t tB = tmem_base + lane_addr + (uint32_t)((1 * 2 + acc) * BN);
                const uint32_t csm = umma::smem_u32(smemCs + (size_t)cst * CS_BYTES);
                // first column of step s inside the N-tile: map row r = s / SPR, this thread's quarter
                auto step_col = [&](int s) -> int { return (D == 128 ? 0 : (s / SPR) * D) + q4 * DQ + (s % SPR) * SW; };
                float vA0[SW], vA1[SW], vB0[SW], vB1[SW];
                float4 c0[SW / 2], c1[SW / 2];
                float hA = 0.f, hB = 0.f;               // halo column (q4 > 0): raw accumulators left of this thread's range
                float4 hc = make_float4(0.f, 0.f, 0.f, 0.f);
                float zprevA = -CUDART_INF_F, zprevB = -CUDART_INF_F;
                float4 obA, obB;
                {
                    const int n0 = step_col(0);
                    umma::tmem_ld_32x8_issue(tA + (uint32_t)n0, vA0);
                    umma::tmem_ld_32x8_issue(tB + (uint32_t)n0, vB0);
#pragma unroll
                    for (int i = 0; i < SW / 2; ++i) c0[i] = umma::lds128(csm + 16 * (n0 / 2 + i));
                    if (q4) {
                        umma::tmem_ld_32x1_issue(tA + (uint32_t)(n0 - 1), hA);
                        umma::tmem_ld_32x1_issue(tB + (uint32_t)(n0 - 1), hB);
                        hc = umma::lds128(csm + 16 * (n0 / 2 - 1));
                    }
                }
#pragma unroll
                for (int s = 0; s < NSTEP; ++s) {
                    float (&vA)[SW] = (s & 1) ? vA1 : vA0;
                    float (&vB)[SW] = (s & 1) ? vB1 : vB0;
                    float (&vAn)[SW] = (s & 1) ? vA0 : vA1;
                    float (&vBn)[SW] = (s & 1) ? vB0 : vB1;
                    float4 (&cc)[SW / 2] = (s & 1) ? c1 : c0;
                    float4 (&cn)[SW / 2] = (s & 1) ? c0 : c1;
                    const int xo = (s % SPR) * SW;          // first column of the step inside this thread's quarter row
                    const int r = (D == 128) ? 0 : s / SPR; // map row inside the N-tile
                    umma::tmem_wait_ld();                   // step s (and its halo) is in registers
                    // halo of THIS step must be consumed before the next step's halo load overwrites it
                    float zhA = -CUDART_INF_F, zhB = -CUDART_INF_F;
                    if (xo == 0 && q4) {
                        zhA = dm_zncc_partial(hA, s1A.x, hc.y, NORMED ? hc.w : 1.0f);
                        zhB = dm_zncc_partial(hB, s1B.x, hc.y, NORMED ? hc.w : 1.0f);
                    }
                    if (s + 1 < NSTEP) {                    // step s+1 in flight during the math below
                        const int n1 = step_col(s + 1);
                        umma::tmem_ld_32x8_issue(tA + (uint32_t)n1, vAn);
                        umma::tmem_ld_32x8_issue(tB + (uint32_t)n1, vBn);
#pragma unroll
                        for (int i = 0; i < SW / 2; ++i) cn[i] = umma::lds128(csm + 16 * (n1 / 2 + i));
                        if (((s + 1) % SPR) == 0 && q4) {
                            umma::tmem_ld_32x1_issue(tA + (uint32_t)(n1 - 1), hA);
                            umma::tmem_ld_32x1_issue(tB + (uint32_t)(n1 - 1), hB);
                            hc = umma::lds128(csm + 16 * (n1 / 2 - 1));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < SW; i += 2) {
                        const float4 cp = cc[i >> 1];       // {s2k0, s2k1, inv0, inv1} of two columns
                        umma::zncc_partial2(vA[i], vA[i + 1], ns1A, cp.x, cp.y, NORMED ? cp.z : 1.0f, NORMED ? cp.w : 1.0f);
                        umma::zncc_partial2(vB[i], vB[i + 1], ns1B, cp.x, cp.y, NORMED ? cp.z : 1.0f, NORMED ? cp.w : 1.0f);
                    }
                    if (MODE == MODE_NULL) { rmaxA = fmaxf(rmaxA, vA[0] + vB[0]); continue; }
                    // map-row parity: compile time when an N-tile holds several map rows; for
                    // D == 128 an N-tile is one map row and the parity alternates with j
                    const bool odd_row = (D == 128) ? ((j & 1) != 0) : ((r & 1) != 0);
#pragma unroll
                    for (int i = 0; i < SW; i += 2) {
                        const int xh = (xo + i) >> 1;       // pooled column inside this thread's quarter
                        const float leftA = (xo + i == 0) ? zhA : (i == 0 ? zprevA : vA[i - 1]);
                        const float leftB = (xo + i == 0) ? zhB : (i == 0 ? zprevB : vB[i - 1]);
                        const float gA = umma::max3(leftA, vA[i], vA[i + 1]);
                        const float gB = umma::max3(leftB, vB[i], vB[i + 1]);
                        rminA = umma::min3(rminA, vA[i], vA[i + 1]);
                        rminB = umma::min3(rminB, vB[i], vB[i + 1]);
                        if (!odd_row) {
                            stA[xh] = fmaxf(stA[xh], gA);   // rows 2y-1 (carried) and 2y
                            stB[xh] = fmaxf(stB[xh], gB);
                        } else {
                            // pooled output: row factor and upper clamp only.  A flat patch has
                            // inv1 = 0 -> the whole row is 0; its min / max are forced to 1 below, so
                            // the slice turns NaN downstream exactly like OpenCV's all-ones map.
                            float oA = fmaxf(stA[xh], gA), oB = fmaxf(stB[xh], gB);
                            if (NORMED) { oA = fminf(__fmul_rn(oA, s1A.y), 1.0f); oB = fminf(__fmul_rn(oB, s1B.y), 1.0f); }
                            stA[xh] = gA; stB[xh] = gB;     // become row 2(y+1)-1 of the next pooled row
                            rmaxA = fmaxf(rmaxA, oA); rmaxB = fmaxf(rmaxB, oB);
                            // index inside the pooled values this N-tile contributes (8; 16 for D == 128)
                            const int ol = (D == 128) ? xh : (r >> 1) * HWQ + xh;
                            if ((ol & 3) == 0) { obA.x = oA; obB.x = oB; } else if ((ol & 3) == 1) { obA.y = oA; obB.y = oB; }
                            else if ((ol & 3) == 2) { obA.z = oA; obB.z = oB; } else { obA.w = oA; obB.w = oB; }
                            if ((ol & 3) == 3) {
                                const int off = lane * 8 + ((((ol >> 2) & 1) ^ wsw) << 2);
                                *reinterpret_cast<float4*>(stgA + off) = obA;
                                *reinterpret_cast<float4*>(stgB + off) = obB;
                            }
                            if ((ol & 7) == 7) {
                                // D = 128: 16 contiguous values of pooled row j/2, flushed in two groups of 8;
                                // D = 64 : 8 contiguous values of pooled row j;
                                // D = 32 : 4 values of pooled row 2j and 4 of row 2j+1
                                constexpr int SEG = HWQ < 8 ? HWQ : 8;
                                const size_t col = (D == 128) ? (size_t)(j >> 1) * (D / 2) + (size_t)q4 * HWQ + (size_t)(ol - 7)
                                                              : (size_t)j * (BN / 4) + (size_t)q4 * HWQ;
                                flush8(stgA, woutA, col, SEG, D / 2);
                                flush8(stgB, woutB, col, SEG, D / 2);
                            }
                        }
                    }
                    zprevA = vA[SW - 1]; zprevB = vB[SW - 1];
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) { umma::mbar_arrive(t_empty + acc); umma::mbar_arrive(c_empty + cst); }
                if (++acc == 2) { acc = 0; accph ^= 1; }
                if (++cst == CS_STAGES) { cst = 0; cph ^= 1; }
            }
            if (MODE == MODE_POOL) {
                // partial min / max of this column quarter; flat patch: OpenCV's map is all ones
                const bool flatA = (s1A.y == 0.0f), flatB = (s1B.y == 0.0f);
                prm.rowmin[4 * prowA + q4] = NORMED ? (flatA ? 1.0f : fminf(fmaxf(__fmul_rn(rminA, s1A.y), -1.0f), 1.0f)) : rminA;
                prm.rowmax[4 * prowA + q4] = (NORMED && flatA) ? 1.0f : rmaxA;
                prm.rowmin[4 * prowB + q4] = NORMED ? (flatB ? 1.0f : fminf(fmaxf(__fmul_rn(rminB, s1B.y), -1.0f), 1.0f)) : rminB;
                prm.rowmax[4 * prowB + q4] = (NORMED && flatB) ? 1.0f : rmaxB;
            }
            if (MODE == MODE_NULL && rmaxA == 12345.678f) prm.pooled[prowA] = rmaxA;     // keep the loads alive
        }
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        umma::tc_fence_after();
        umma::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int MODE, int D, bool NORMED>
int launch2(const CUtensorMap& mapA, const CUtensorMap& mapB, const Params& prm, cudaStream_t stream) {
    static bool configured = false;
    auto kern = dm_correlation_p4_kernel<MODE, D, NORMED>;
    if (!configured) {
        DM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured = true;
    }
    int dev = 0, sms = 0;
    DM_CUDA_CHECK(cudaGetDevice(&dev));
    DM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = prm.n_items < sms ? prm.n_items : sms;
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(mapA, mapB, prm);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

template <int MODE, int D>
int launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const Params& prm, bool normed, cudaStream_t stream) {
    return normed ? launch2<MODE, D, true>(mapA, mapB, prm, stream) : launch2<MODE, D, false>(mapA, mapB, prm, stream);
}

}  // namespace

bool dm_correlation_p4_pool_supported(int t0, int t1, int kpad) {
    const int p = t0 * t1;
    return p >= HALVES * BM && p % (HALVES * BM) == 0 && kpad % BK == 0 && kpad <= MAX_KB * BK &&
           (t1 == 32 || t1 == 64 || t1 == 128) && t0 % 2 == 0;
}

int dm_correlation_p4_pool(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                           int n_tiles, int t0, int t1, int kpad, int kreal, int method, int null_epilogue,
                           float* pooled, float* rowmin, float* rowmax, cudaStream_t stream) {
    DM_REQUIRE(dm_correlation_p4_pool_supported(t0, t1, kpad), DM_ERR_UNSUPPORTED, "16-warp pooled tcgen05 correlation: unsupported grid (%d,%d)", t0, t1);
    const int p = t0 * t1;
    const uint64_t rows = (uint64_t)n_tiles * p;
    Params prm; CUtensorMap mapA, mapB;
    int rc = dm_make_desc_tensor_map(&mapA, desc1, rows, kpad, BM);
    if (rc != DM_OK) return rc;
    rc = dm_make_desc_tensor_map(&mapB, desc2, rows, kpad, BN);
    if (rc != DM_OK) return rc;
    prm.stat1 = (const dm_stat*)stat1;
    prm.cstat2 = reinterpret_cast<const float4*>((const dm_stat*)stat2 + rows);
    prm.P = p; prm.KB = kpad / BK; prm.items_per_tile = p / (HALVES * BM);
    prm.ksteps = (kreal + UMMA_K - 1) / UMMA_K;
    if (prm.ksteps > kpad / UMMA_K || prm.ksteps <= 0) prm.ksteps = kpad / UMMA_K;
    prm.n_items = n_tiles * prm.items_per_tile;
    prm.pooled = pooled; prm.rowmin = rowmin; prm.rowmax = rowmax;
    const bool normed = method == DM_TM_CCOEFF_NORMED;
    if (null_epilogue) return launch<MODE_NULL, 64>(mapA, mapB, prm, true, stream);
    if (t1 == 128) return launch<MODE_POOL, 128>(mapA, mapB, prm, normed, stream);
    if (t1 == 64) return launch<MODE_POOL, 64>(mapA, mapB, prm, normed, stream);
    return launch<MODE_POOL, 32>(mapA, mapB, prm, normed, stream);
}
