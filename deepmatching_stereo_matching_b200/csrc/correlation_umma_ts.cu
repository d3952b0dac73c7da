// tcgen05 correlation engine with the patch operand RESIDENT IN TENSOR MEMORY (sm_100a).
//
// Same contraction and the same pooled ZNCC epilogue as correlation_umma.cu (MODE_POOL):
// replaces the P calls of cv2.matchTemplate in Correlation_map._create_simple_initial_co_map
// (misc/Correlation_map.py:69-87) fused with the 3x3/s2/p1 max-pool of
// Correlation_map._aggregation (misc/Correlation_map.py:100-103) and the per-patch
// min / max of Feature_value.min_max (misc/Feature_value.py:32-37).
//
// Why a second engine: ncu showed the shared-memory data pipe of the SS kernel 100 % busy
// (tensor-core operand fetch 47 % + LDS/STS 30 % + tcgen05.ld 23 % of the cycles).  Half of
// the operand fetch is the A block (the item's 256 patch descriptors), which never changes
// during the sweep over the tile's positions.  Here A is copied ONCE per item from shared
// memory into TMEM (tcgen05.st) and every MMA takes it from there
// (tcgen05.mma [d], [a_tmem], b_desc), which halves the operand traffic.  K steps that hold
// only padding are skipped (ws = 15: 15 of 16).
//
// TMEM (512 columns):  [0,128) A half 0   [128,256) A half 1   (bf16 pairs, K <= 256)
//                      [256,512) accumulators: (half h, stage s) at 256 + (2 h + s) * 64
// Work item = 256 patches of one tile x all P positions, swept in N-tiles of 64 positions.
// Persistent grid, one CTA per SM, 12 warps in three warpgroups (setmaxnreg 40 / 232):
//   warp 0      TMA producer   A boxes of the NEXT item (landing zone in shared memory),
//                              ring of B boxes (64 positions x 64 bf16, 128-byte swizzle),
//                              ring of 512-byte column tables
//   warp 1      MMA issuer     tcgen05.mma cta_group::1 kind::f16, M=128 N=64 K=16, A from TMEM
//   warps 4..11 epilogue       as in correlation_umma.cu: a thread owns two patch rows (TMEM
//                              lane l of both halves) and half of the columns of every map
//                              row; 8-column software-pipelined steps; results leave through
//                              an XOR-swizzled (conflict-free) shared-memory transpose so that
//                              every store instruction writes 8 rows x 64 B.  The epilogue
//                              warps also move the next item's A block into TMEM as soon as
//                              the last accumulator of the current item is complete.
#include "dm_common.cuh"
#include "dm_internal.h"
#include "umma.cuh"

namespace {

constexpr int BM = 128;                 // rows per accumulator half = TMEM lanes
constexpr int HALVES = 2;
constexpr int BN = 64;                  // positions per N-tile
constexpr int BK = 64;                  // bf16 per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 6;               // B ring
constexpr int MAX_KB = 4;               // kpad <= 256
constexpr int A_BOX_BYTES = BM * BK * 2;    // 16 KiB
constexpr int B_BOX_BYTES = BN * BK * 2;    // 8 KiB
constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int TMEM_COLS = 512;
constexpr int TM_A = 0;                 // A half h at TM_A + h * 128
constexpr int TM_ACC = 256;             // accumulator (h, stage) at TM_ACC + (2 h + stage) * BN
constexpr int SW = 8;                   // columns per epilogue pipeline step
constexpr int NSTEP = (BN / 2) / SW;    // 4: a thread covers half of the N-tile's columns
constexpr int STG_FLOATS = 32 * 16;     // one staging region: 32 rows x 16 floats, XOR swizzled
constexpr int STG_BYTES = 2 * STG_FLOATS * 4;   // per epilogue warp: one region per accumulator half
constexpr int CS_BYTES = (BN / 2) * 16; // column table of one N-tile: 32 x {sk0, sk1, inv0, inv1}
constexpr int CS_STAGES = 8;

constexpr size_t OFF_A = 0;
constexpr size_t OFF_B = OFF_A + (size_t)HALVES * MAX_KB * A_BOX_BYTES;
constexpr size_t OFF_STG = OFF_B + (size_t)STAGES * B_BOX_BYTES;
constexpr size_t OFF_CS = OFF_STG + (size_t)EPI_WARPS * STG_BYTES;
constexpr size_t OFF_BAR = OFF_CS + CS_STAGES * CS_BYTES;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + OFF_BAR + 512;

struct Params {
    const dm_stat* stat1;       // [n*P] float4 {S', inv, S'/K, mean}
    const float4* cstat2;       // [n*P/2] {S'/K even, S'/K odd, inv even, inv odd} of image 2
    int n_items, P, KB, ksteps, items_per_tile;
    float* pooled;              // [n][P][P/4]
    float* rowmin; float* rowmax;   // [n][P][4] partial min / max: each column half writes its value twice (16-warp kernel: quarters)
};

enum { MODE_POOL = 1, MODE_NULL = 2 };   // MODE_NULL: drain TMEM only (measurement aid)

template <int MODE, int D, bool NORMED>      // D = positions per map row (T1): 16, 32 or 64
__global__ void __launch_bounds__(THREADS, 1)
dm_correlation_ts_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smemA = smem + OFF_A;
    uint8_t* smemB = smem + OFF_B;
    float* smemStg = reinterpret_cast<float*>(smem + OFF_STG);
    uint8_t* smemCs = smem + OFF_CS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;                    // TMA: A block landed in shared memory
    uint64_t* a_empty = bars + 1;               // epilogue: A block copied out of shared memory
    uint64_t* at_full = bars + 2;               // epilogue: A block is in TMEM
    uint64_t* b_full = bars + 3;
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
        umma::mbar_init(a_empty, EPI_WARPS);
        umma::mbar_init(at_full, EPI_WARPS);
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
            auto load_a = [&](int item) {
                const int tile = item / prm.items_per_tile;
                const int row0 = tile * P + (item - tile * prm.items_per_tile) * (HALVES * BM);
                umma::mbar_wait(a_empty, aph ^ 1);          // landing zone free (previous block copied to TMEM)
                aph ^= 1;
                if (umma::elect_one()) {
                    umma::mbar_expect_tx(a_full, (uint32_t)(HALVES * KB * A_BOX_BYTES));
                    for (int h = 0; h < HALVES; ++h)
                        for (int kb = 0; kb < KB; ++kb)
                            umma::tma_load_2d(smemA + (size_t)(h * MAX_KB + kb) * A_BOX_BYTES, &mapA, kb * BK, row0 + h * BM, a_full);
                }
                __syncwarp();
            };
            if ((int)blockIdx.x < prm.n_items) load_a(blockIdx.x);
            for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
                const int tile = item / prm.items_per_tile;
                for (int j = 0; j < NT; ++j) {
                    for (int kb = 0; kb < KB; ++kb) {
                        umma::mbar_wait(b_empty + bs, bph ^ 1);
                        if (umma::elect_one()) {
                            umma::mbar_expect_tx(b_full + bs, (uint32_t)B_BOX_BYTES);
                            umma::tma_load_2d(smemB + (size_t)bs * B_BOX_BYTES, &mapB, kb * BK, tile * P + j * BN, b_full + bs);
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
                    // the next item's A block, one item ahead (its landing zone frees early in this item)
                    if (j == 1 && item + (int)gridDim.x < prm.n_items) load_a(item + gridDim.x);
                }
            }
        } else if (warp == 1) {
            // -------------------------------------------------------- MMA issuer (converged warp, elected issue)
            constexpr uint32_t idesc = umma::instr_desc_bf16(BM, BN);
            constexpr int KPB = BK / UMMA_K;                // K steps per 64-wide k-block
            int bs = 0; uint32_t bph = 0, atph = 0; int acc = 0; uint32_t accph = 0;
            for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
                umma::mbar_wait(at_full, atph);             // this item's A block is in TMEM
                atph ^= 1;
                umma::tc_fence_after();
                for (int j = 0; j < NT; ++j) {
                    umma::mbar_wait(t_empty + acc, accph ^ 1);
                    umma::tc_fence_after();
                    const uint32_t d0 = tmem_base + (uint32_t)(TM_ACC + acc * BN);
                    for (int kb = 0; kb < KB; ++kb) {
                        umma::mbar_wait(b_full + bs, bph);
                        umma::tc_fence_after();
                        const uint64_t bdesc = umma::smem_desc_sw128(smemB + (size_t)bs * B_BOX_BYTES);
                        const uint32_t a0 = tmem_base + (uint32_t)(TM_A + kb * KPB * (UMMA_K / 2));
                        const int nk = prm.ksteps - kb * KPB;   // padding-only K steps are skipped
#pragma unroll
                        for (int k = 0; k < KPB; ++k) {
                            if (k < nk) {
#pragma unroll
                                for (int h = 0; h < HALVES; ++h)
                                    if (umma::elect_one())
                                        umma::mma_bf16_ts(d0 + (uint32_t)(h * 2 * BN), a0 + (uint32_t)(h * 128 + k * (UMMA_K / 2)),
                                                          bdesc + 2 * k, idesc, (kb | k) != 0);
                            }
                        }
                        if (umma::elect_one()) umma::mma_commit(b_empty + bs);
                        if (++bs == STAGES) { bs = 0; bph ^= 1; }
                    }
                    if (umma::elect_one()) umma::mma_commit(t_full + acc);
                    if (++acc == 2) { acc = 0; accph ^= 1; }
                }
            }
        }
    } else {
        umma::reg_alloc<232>();
        // ------------------------------------------------------------ epilogue
        const int e = warp - FIRST_EPI_WARP;
        const int quarter = warp & 3;                   // TMEM lanes a warp may touch: 32*(warp_id % 4)..+31
        const int ch = e >> 2;                          // which half of every map row's columns
        constexpr int DH = D / 2;                       // columns of a map row handled by this thread
        constexpr int HWQ = D / 4;                      // pooled outputs per map row per thread
        constexpr int RPT = BN / D;                     // map rows per N-tile
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        float* stgA = smemStg + (size_t)e * (2 * STG_FLOATS);
        float* stgB = stgA + STG_FLOATS;
        const int wsw = (lane >> 1) & 3;                // XOR swizzle of this lane's own staging row
        int acc = 0; uint32_t accph = 0; int cst = 0; uint32_t cph = 0; uint32_t aph = 0;

        // A block: shared memory (128-byte swizzled TMA boxes) -> TMEM, 16 columns per k-block
        // and column half; row l of the box sits at l*128 with its 16-byte chunks XORed by (l & 7)
        auto copy_a = [&]() {
            umma::mbar_wait(a_full, aph);
            aph ^= 1;
            const int rib = quarter * 32 + lane;        // row inside the box = TMEM lane
#pragma unroll
            for (int h = 0; h < HALVES; ++h) {
                for (int kb = 0; kb < KB; ++kb) {
                    const uint32_t rowaddr = umma::smem_u32(smemA + (size_t)(h * MAX_KB + kb) * A_BOX_BYTES + (size_t)rib * 128);
                    uint32_t v[16];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 q = umma::lds128u(rowaddr + (uint32_t)(((ch * 4 + c) ^ (rib & 7)) << 4));
                        v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
                    }
                    umma::tmem_st_32x16(tmem_base + lane_addr + (uint32_t)(TM_A + h * 128 + kb * 32 + ch * 16), v);
                }
            }
            umma::tmem_wait_st();
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) { umma::mbar_arrive(at_full); umma::mbar_arrive(a_empty); }
        };

        if ((int)blockIdx.x < prm.n_items) copy_a();
        for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
            const int tile = item / prm.items_per_tile;
            const size_t wrow = (size_t)tile * P + (size_t)(item - tile * prm.items_per_tile) * (HALVES * BM) + quarter * 32;
            const size_t prowA = wrow + lane, prowB = prowA + BM;
            const dm_stat s1A = prm.stat1[prowA], s1B = prm.stat1[prowB];
            const bool flatA = (s1A.y == 0.0f), flatB = (s1B.y == 0.0f);
            const float ns1A = -s1A.x, ns1B = -s1B.x;
            // st[] = horizontally pooled previous map row (odd rows) / running vertical max (even rows)
            float stA[HWQ], stB[HWQ];
            float rminA = CUDART_INF_F, rmaxA = -CUDART_INF_F, rminB = CUDART_INF_F, rmaxB = -CUDART_INF_F;
#pragma unroll
            for (int i = 0; i < HWQ; ++i) { stA[i] = -CUDART_INF_F; stB[i] = -CUDART_INF_F; }
            const size_t ostride = (size_t)(P / 4);
            float* woutA = prm.pooled + wrow * ostride;
            float* woutB = woutA + (size_t)BM * ostride;
            // 16 staged floats per lane -> the warp writes 8 rows x 64 B per store instruction.
            // The 16 floats are segments of `seg` floats, `segstride` floats apart in the output.
            auto flush16 = [&](const float* stg, float* wout, size_t col, int seg, size_t segstride) {
                __syncwarp();
                const int f = lane & 3, f4 = f * 4;
                const uint32_t sa = umma::smem_u32(stg + (lane >> 2) * 16 + ((f ^ ((lane >> 3) & 3)) << 2));
                float4 v[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) v[it] = umma::lds128(sa + (uint32_t)(it * 8 * 16 * 4));
                float* dst = wout + (size_t)(lane >> 2) * ostride + col + (size_t)(f4 / seg) * segstride + (f4 % seg);
#pragma unroll
                for (int it = 0; it < 4; ++it) *reinterpret_cast<float4*>(dst + (size_t)(it * 8) * ostride) = v[it];
                __syncwarp();
            };
            for (int j = 0; j < NT; ++j) {
                umma::mbar_wait(c_full + cst, cph);
                umma::mbar_wait(t_full + acc, accph);
                umma::tc_fence_after();
                // every MMA of this item has completed: the next item's A block may replace it
                if (j == NT - 1 && item + (int)gridDim.x < prm.n_items) copy_a();
                const uint32_t tA = tmem_base + lane_addr + (uint32_t)(TM_ACC + (0 * 2 + acc) * BN);
                const uint32_t tB = tmem_base + lane_addr + (uint32_t)(TM_ACC + (1 * 2 + acc) * BN);
                const uint32_t csm = umma::smem_u32(smemCs + (size_t)cst * CS_BYTES);
                // column of step s inside the N-tile: map row r = (8 s) / DH, then this thread's half
                auto step_col = [&](int s) -> int { return ((s * SW) / DH) * D + ch * DH + (s * SW) % DH; };
                float vA0[SW], vA1[SW], vB0[SW], vB1[SW];
                float4 c0[SW / 2], c1[SW / 2];
                float hA = 0.f, hB = 0.f;               // halo column (ch == 1): raw accumulators left of the split
                float4 hc = make_float4(0.f, 0.f, 0.f, 0.f);
                float zprevA = -CUDART_INF_F, zprevB = -CUDART_INF_F;
                float4 obA, obB;
                {
                    const int n0 = step_col(0);
                    umma::tmem_ld_32x8_issue(tA + (uint32_t)n0, vA0);
                    umma::tmem_ld_32x8_issue(tB + (uint32_t)n0, vB0);
#pragma unroll
                    for (int i = 0; i < SW / 2; ++i) c0[i] = umma::lds128(csm + 16 * (n0 / 2 + i));
                    if (ch) {
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
                    const int xo = (s * SW) % DH;           // first column of the step inside this thread's row half
                    const int r = (s * SW) / DH;            // map row inside the N-tile
                    umma::tmem_wait_ld();                   // step s (and its halo) is in registers
                    // halo of THIS step must be consumed before the next step's halo load overwrites it
                    float zhA = -CUDART_INF_F, zhB = -CUDART_INF_F;
                    if (xo == 0 && ch) {
                        zhA = dm_zncc_partial(hA, s1A.x, hc.y, NORMED ? hc.w : 1.0f);
                        zhB = dm_zncc_partial(hB, s1B.x, hc.y, NORMED ? hc.w : 1.0f);
                    }
                    if (s + 1 < NSTEP) {                    // step s+1 in flight during the math below
                        const int n1 = step_col(s + 1);
                        umma::tmem_ld_32x8_issue(tA + (uint32_t)n1, vAn);
                        umma::tmem_ld_32x8_issue(tB + (uint32_t)n1, vBn);
#pragma unroll
                        for (int i = 0; i < SW / 2; ++i) cn[i] = umma::lds128(csm + 16 * (n1 / 2 + i));
                        if (((s + 1) * SW) % DH == 0 && ch) {
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
                    // map-row parity: an N-tile holds RPT = 64 / D map rows; for D == 64 it is one
                    // map row and the parity alternates with j
                    const bool odd_row = (RPT == 1) ? ((j & 1) != 0) : ((r & 1) != 0);
                    if (xo == 0) { zprevA = -CUDART_INF_F; zprevB = -CUDART_INF_F; }
#pragma unroll
                    for (int i = 0; i < SW; i += 2) {
                        const int xh = (xo + i) >> 1;       // pooled column inside this thread's half
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
                            // position inside the 8 outputs this N-tile contributes (16 for D == 64,
                            // where only odd N-tiles produce output)
                            const int ol = (RPT == 1) ? xh : (r >> 1) * HWQ + xh;
                            if ((ol & 3) == 0) { obA.x = oA; obB.x = oB; } else if ((ol & 3) == 1) { obA.y = oA; obB.y = oB; }
                            else if ((ol & 3) == 2) { obA.z = oA; obB.z = oB; } else { obA.w = oA; obB.w = oB; }
                            if ((ol & 3) == 3) {
                                const int slot = ((RPT == 1) ? 0 : (j & 1) * 2) + (ol >> 2);     // float4 slot 0..3 of the 16
                                const int off = lane * 16 + ((slot ^ wsw) << 2);
                                *reinterpret_cast<float4*>(stgA + off) = obA;
                                *reinterpret_cast<float4*>(stgB + off) = obB;
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
                if (MODE == MODE_POOL && (j & 1)) {
                    // 16 pooled values per row are staged: the pooled rows completed by N-tiles j-1 and j
                    constexpr int SEG = HWQ < 16 ? HWQ : 16;
                    const size_t col = (size_t)(j >> 1) * 32 + (size_t)ch * HWQ;
                    flush16(stgA, woutA, col, SEG, DH);
                    flush16(stgB, woutB, col, SEG, DH);
                }
            }
            if (MODE == MODE_POOL) {
                // partial min / max of this column half; flat patch: OpenCV's map is all ones
                { const float v = NORMED ? (flatA ? 1.0f : fminf(fmaxf(__fmul_rn(rminA, s1A.y), -1.0f), 1.0f)) : rminA; prm.rowmin[4 * prowA + 2 * ch] = v; prm.rowmin[4 * prowA + 2 * ch + 1] = v; }
                { const float v = (NORMED && flatA) ? 1.0f : rmaxA; prm.rowmax[4 * prowA + 2 * ch] = v; prm.rowmax[4 * prowA + 2 * ch + 1] = v; }
                { const float v = NORMED ? (flatB ? 1.0f : fminf(fmaxf(__fmul_rn(rminB, s1B.y), -1.0f), 1.0f)) : rminB; prm.rowmin[4 * prowB + 2 * ch] = v; prm.rowmin[4 * prowB + 2 * ch + 1] = v; }
                { const float v = (NORMED && flatB) ? 1.0f : rmaxB; prm.rowmax[4 * prowB + 2 * ch] = v; prm.rowmax[4 * prowB + 2 * ch + 1] = v; }
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
    auto kern = dm_correlation_ts_kernel<MODE, D, NORMED>;
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

bool dm_correlation_ts_pool_supported(int t0, int t1, int kpad) {
    const int p = t0 * t1;
    return p >= HALVES * BM && p % (HALVES * BM) == 0 && kpad % BK == 0 && kpad <= MAX_KB * BK &&
           (t1 == 16 || t1 == 32 || t1 == 64) && t0 % 2 == 0 && (p / BN) % 2 == 0;
}

int dm_correlation_ts_pool(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                           int n_tiles, int t0, int t1, int kpad, int kreal, int method, int null_epilogue,
                           float* pooled, float* rowmin, float* rowmax, cudaStream_t stream) {
    DM_REQUIRE(dm_correlation_ts_pool_supported(t0, t1, kpad), DM_ERR_UNSUPPORTED, "TMEM-resident tcgen05 correlation: unsupported grid (%d,%d)", t0, t1);
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
    if (t1 == 64) return launch<MODE_POOL, 64>(mapA, mapB, prm, normed, stream);
    if (t1 == 32) return launch<MODE_POOL, 32>(mapA, mapB, prm, normed, stream);
    return launch<MODE_POOL, 16>(mapA, mapB, prm, normed, stream);
}
