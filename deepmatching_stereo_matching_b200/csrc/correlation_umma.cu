// tcgen05 / TMEM / TMA correlation engine (sm_100a).
//
// Replaces the P calls of cv2.matchTemplate in
// Correlation_map._create_simple_initial_co_map (misc/Correlation_map.py:69-87):
// per tile the dense contraction C[p][q] = sum_k A1[p][k] * A2[q][k] over the bf16
// descriptor matrices (exact integers, fp32 accumulate exact), followed by the ZNCC
// epilogue.  Two epilogues:
//   MODE_RAW  : writes raw ZNCC  [tile][p][q]                    (materialising path)
//   MODE_POOL : 3x3/s2/p1 max-pools the raw ZNCC of every patch on the fly (the first
//               half of Correlation_map._aggregation, misc/Correlation_map.py:100-103;
//               min-max and **1.4 are monotone, so pooling commutes with them) and writes
//               only the pooled map [tile][p][C/2][D/2] plus the per-patch min / max that
//               Feature_value.min_max (misc/Feature_value.py:32-37) needs.  Level 0 never
//               reaches HBM.
//
// Work item = 256 patches of one tile (two M=128 accumulator halves) x all P positions,
// swept in N-tiles of 128 positions.  Persistent grid, one CTA per SM, 12 warps in three
// warpgroups (setmaxnreg gives the producer/MMA group 40 registers and the epilogue 232):
//   warp 0      TMA producer   A (2 halves x KB boxes, once per item), B ring of 3 stages,
//                              and the N-tile's 1 KiB column table (bulk copy) into a
//                              4-deep shared-memory ring
//   warp 1      MMA issuer     tcgen05.mma cta_group::1 kind::f16, M=128 N=128 K=16;
//                              TMEM: 2 halves x 2 accumulator stages x 128 columns = 512
//   warps 4..11 epilogue       tcgen05.ld 32x32b.  A thread owns TWO patch rows -- TMEM lane l
//                              of both accumulator halves -- and one half of the columns
//                              of every map row, so the 3x3 pooling, the running row
//                              minimum and the halo row carried between N-tiles are
//                              thread-local registers, the per-column parameters are
//                              fetched once for two rows (the shared-memory broadcast is
//                              what the epilogue pays most for) and two independent
//                              dependency chains interleave.  The one halo column at the
//                              split comes from a 1-column TMEM load.  Software pipelined
//                              in steps of 8 columns (loads of step s+1 in flight during
//                              step s), FFMA2/FMUL2 packed math, 3-input min/max.  Pooled
//                              results are staged in shared memory as regions of 32 rows x
//                              128 B (the layout of a 128-byte-swizzled TMA box; at D = 64 / 32
//                              the two column halves of a row meet there, staged by the warp
//                              pair e, e + 4) and leave by 2-D bulk tensor stores; the raw mode
//                              and D = 16 keep a per-warp transpose + LDS/STG flush (8 rows x
//                              64 B per store instruction).
// B traffic: every CTA streams the tile's whole position matrix once per item; with
// M = 256 rows per item that is 32 B/clk/SM from L2 at the MMA's full rate.
#include <atomic>
#include <cstdio>
#include <cstdlib>
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
constexpr int STAGES_PAIR = 6;          // CTA-pair kernel: a B stage holds 64 of the N-tile's 128 rows (8 KiB)
constexpr int MAX_KB = 4;               // kpad <= 256
constexpr int BOX_BYTES = BM * BK * 2;  // 16 KiB: one [128 rows x 64] bf16 box
constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;        // warps 0..3 = warpgroup 0: TMA producer, MMA issuer, 2 idle
constexpr int THREADS = 32 * (FIRST_EPI_WARP + EPI_WARPS);   // 3 warpgroups
constexpr int TMEM_COLS = 512;
constexpr int SW = 8;                   // columns per epilogue pipeline step (per accumulator half)
constexpr int NSTEP = (BN / 2) / SW;    // a thread covers half of the N-tile's columns
constexpr int STG_STRIDE = 16;          // floats per staged row; float4 slots XOR-swizzled by (row >> 1) & 3: conflict-free both ways
constexpr int STG_BYTES = 2 * 32 * STG_STRIDE * 4;  // per epilogue warp: one region per accumulator half
constexpr int CS_BYTES = BN * 4;        // column table of one N-tile: inv2 of its 128 positions (the S'/K term rides in the descriptors' correction slots)
constexpr int CS_STAGES = 4;

// shared memory carve-up (offsets from the 1024-aligned base)
constexpr size_t OFF_A = 0;
constexpr size_t OFF_B = OFF_A + (size_t)HALVES * MAX_KB * BOX_BYTES;
constexpr size_t OFF_STG = OFF_B + (size_t)STAGES * BOX_BYTES;
constexpr size_t OFF_CS = OFF_STG + (size_t)EPI_WARPS * STG_BYTES;
constexpr size_t OFF_BAR = OFF_CS + CS_STAGES * CS_BYTES;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + OFF_BAR + 256;

// Per-instance carve-up.  WIDE (pooled epilogue at D == 128): a thread completes 32 consecutive pooled floats of a
// patch row per map-row pair -- a whole 128-byte line -- but 16-float staging rows made it leave as two 64-byte
// pieces (8 rows x 64 B per store instruction: 8 L1 wavefronts for 512 bytes).  Staging rows of 32 floats need
// 8 KiB per epilogue warp; the 32 KiB come from the B ring (4 stages of 8 KiB = one N-tile of K blocks on CTA
// pairs, 2 of 16 KiB on a single CTA) and from the column-table ring (2 stages).
template <bool WIDE, bool PAIR>
struct Layout {
    static constexpr int NSTG = PAIR ? (WIDE ? 4 : STAGES_PAIR) : (WIDE ? 2 : STAGES);
    static constexpr int B_STAGE = PAIR ? BOX_BYTES / 2 : BOX_BYTES;
    static constexpr int CSN = WIDE ? 2 : CS_STAGES;
    static constexpr int STG_WARP = WIDE ? 2 * STG_BYTES : STG_BYTES;
    static constexpr size_t O_STG = OFF_B + (size_t)NSTG * B_STAGE;
    static constexpr size_t O_CS = O_STG + (size_t)EPI_WARPS * STG_WARP;
    static constexpr size_t O_BAR = O_CS + (size_t)CSN * CS_BYTES;
    static constexpr size_t SMEM = 1024 /*align slack*/ + O_BAR + 256;
};
static_assert(Layout<false, true>::SMEM == SMEM_BYTES && Layout<false, false>::SMEM == SMEM_BYTES, "the default carve-up is the one above");
static_assert(Layout<true, true>::SMEM <= 227 * 1024 && Layout<true, false>::SMEM <= 227 * 1024, "wide staging does not fit");

struct Params {
    const dm_stat* stat1;       // [n*P] float4 {S', inv, S'/K, mean}
    const float* inv2;          // [n*P] inv of the windows of image 2 (compact table written by dm_descriptors)
    int n_items, P, KB, ksteps, items_per_tile;
    float* raw;                 // MODE_RAW : [n][P][P]
    float* pooled;              // MODE_POOL: [n][P][P/4]
    float* rowmin; float* rowmax;   // MODE_POOL: [n][P][4] partial min / max: each column half writes its value twice
    int pair_flush;                 // MODE_POOL: 0 per-warp flush, 1 pair flush (D == 64 / 32), 2 + tensor stores (also D == 128): the two column halves of a patch row leave as one 128-byte line (below)
};

enum { MODE_RAW = 0, MODE_POOL = 1, MODE_NULL = 2 };   // MODE_NULL: drain TMEM only (measurement aid)

// PAIR: the CTA pair of a 2-CTA cluster works on 512 patches of one tile (256 per CTA, epilogue
// unchanged) with tcgen05.mma.cta_group::2 (M = 256 across the pair, N = 128): each CTA fetches
// only HALF of every B tile from its own shared memory and streams only half of B from L2.
template <int MODE, int D, bool NORMED, bool PAIR, bool WIDE>      // D = positions per map row (T1); only used by MODE_POOL
__global__ void __launch_bounds__(THREADS, 1)
dm_correlation_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                           const __grid_constant__ CUtensorMap mapO, const Params prm) {
    extern __shared__ uint8_t smem_raw[];
    // keep shared-space provenance: offset the array instead of round-tripping through integers
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smemA = smem + OFF_A;
    uint8_t* smemB = smem + OFF_B;
    using L = Layout<WIDE, PAIR>;
    static_assert(!WIDE || (MODE == MODE_POOL && D == 128), "wide staging exists for the pooled epilogue at D == 128");
    float* smemStg = reinterpret_cast<float*>(smem + L::O_STG);
    uint8_t* smemCs = smem + L::O_CS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::O_BAR);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + 1;
    constexpr int NSTG = L::NSTG;
    constexpr int CSN = L::CSN;
    constexpr int ISSUERS = 1;              // 2: warps 1 and 2 of the leader each issue the MMAs of one accumulator half
    constexpr int B_STAGE_BYTES = PAIR ? BOX_BYTES / 2 : BOX_BYTES;
    uint64_t* b_full = bars + 2;
    uint64_t* b_empty = bars + 2 + NSTG;
    uint64_t* t_full = bars + 2 + 2 * NSTG;
    uint64_t* t_empty = bars + 4 + 2 * NSTG;
    uint64_t* c_full = bars + 6 + 2 * NSTG;
    uint64_t* c_empty = bars + 6 + 2 * NSTG + CSN;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6 + 2 * NSTG + 2 * CSN);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int P = prm.P, KB = prm.KB, NT = P / BN;
    // work slots: a slot is a CTA (or a CTA pair); the pair's work unit is two consecutive items of one tile
    // (shuffled so that ptxas can prove the value warp-uniform: otherwise every MMA / TMA issue behind
    //  `rank == 0` is wrapped in an ELECT / BRA.U.ANY loop, ~150 cycles per MMA)
    //  the cluster is (2,1,1) on a 1-D grid, so %cluster_ctarank == blockIdx.x & 1
    const int rank = PAIR ? (int)(blockIdx.x & 1u) : 0;
    const int slot = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_slots = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_units = PAIR ? prm.n_items >> 1 : prm.n_items;

    if (threadIdx.x == 0) {
        umma::mbar_init(a_full, 1);
        umma::mbar_init(a_empty, ISSUERS);
        for (int s = 0; s < NSTG; ++s) { umma::mbar_init(b_full + s, 1); umma::mbar_init(b_empty + s, ISSUERS); }
        for (int s = 0; s < 2; ++s) { umma::mbar_init(t_full + s, ISSUERS); umma::mbar_init(t_empty + s, (PAIR ? 2 : 1) * EPI_WARPS); }
        for (int s = 0; s < CSN; ++s) { umma::mbar_init(c_full + s, 1); umma::mbar_init(c_empty + s, EPI_WARPS); }
        umma::fence_barrier_init();
        umma::tma_prefetch_desc(&mapA);
        umma::tma_prefetch_desc(&mapB);
    }
    if (warp == 1) {
        if (PAIR) { umma::tmem_alloc_pair(tmem_slot, TMEM_COLS); umma::tmem_relinquish_pair(); }
        else { umma::tmem_alloc(tmem_slot, TMEM_COLS); umma::tmem_relinquish(); }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (PAIR) umma::cluster_sync_all();     // the peer's barriers are initialised before anything signals them
    umma::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // programmatic dependent launch: everything above (barriers, TMEM, tensor-map prefetch, cluster sync) ran
    // while the descriptor kernel was still draining; its rows and statistics are read only from here on
    dm_pdl_wait();
    dm_pdl_launch_dependents();

    // register re-partitioning between the warpgroups (setmaxnreg): the producer / MMA
    // warpgroup keeps 40 registers per thread, the two epilogue warpgroups get 232
    if (warp < FIRST_EPI_WARP) {
    umma::reg_dealloc<40>();
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (converged warp, elected issue)
        // PAIR: both CTAs load into their own shared memory; the bytes of both are counted by the
        // LEADER's full barriers (the MMA issuer waits there), the empty barriers are local
        // (the leader's commits are multicast to both CTAs)
        int bs = 0; uint32_t bph = 0, aph = 0; int cst = 0; uint32_t cph = 0;
        const uint32_t a_full_lead = PAIR ? umma::mapa_u32(a_full, 0) : 0;
        const uint32_t b_full_lead = PAIR ? umma::mapa_u32(b_full, 0) : 0;
        for (int unit = slot; unit < n_units; unit += n_slots) {
            const int item = PAIR ? unit * 2 + rank : unit;
            const int tile = item / prm.items_per_tile;
            const int row0 = tile * P + (item - tile * prm.items_per_tile) * (HALVES * BM);
            umma::mbar_wait(a_empty, aph ^ 1);
            if (umma::elect_one()) {
                if (!PAIR || rank == 0) umma::mbar_expect_tx(a_full, (uint32_t)((PAIR ? 2 : 1) * HALVES * KB * BOX_BYTES));
                for (int h = 0; h < HALVES; ++h)
                    for (int kb = 0; kb < KB; ++kb) {
                        if (PAIR) umma::tma_load_2d_pair(smemA + (size_t)(h * MAX_KB + kb) * BOX_BYTES, &mapA, kb * BK, row0 + h * BM, a_full_lead);
                        else umma::tma_load_2d(smemA + (size_t)(h * MAX_KB + kb) * BOX_BYTES, &mapA, kb * BK, row0 + h * BM, a_full);
                    }
            }
            __syncwarp();
            aph ^= 1;
            for (int j = 0; j < NT; ++j) {
                for (int kb = 0; kb < KB; ++kb) {
                    umma::mbar_wait(b_empty + bs, bph ^ 1);
                    if (umma::elect_one()) {
                        if (PAIR) {     // this CTA's half of the N-tile's rows: rank 0 supplies columns 0..63 of the MMA, rank 1 columns 64..127
                            if (rank == 0) umma::mbar_expect_tx(b_full + bs, (uint32_t)(2 * B_STAGE_BYTES));
                            umma::tma_load_2d_pair(smemB + (size_t)bs * B_STAGE_BYTES, &mapB, kb * BK, tile * P + j * BN + rank * (BN / 2),
                                                   b_full_lead + (uint32_t)(bs * 8));
                        } else {
                            umma::mbar_expect_tx(b_full + bs, (uint32_t)BOX_BYTES);
                            umma::tma_load_2d(smemB + (size_t)bs * BOX_BYTES, &mapB, kb * BK, tile * P + j * BN, b_full + bs);
                        }
                    }
                    __syncwarp();
                    if (++bs == NSTG) { bs = 0; bph ^= 1; }
                }
                // column table of this N-tile (needed only by the epilogue, after the MMAs)
                umma::mbar_wait(c_empty + cst, cph ^ 1);
                if (umma::elect_one()) {
                    umma::mbar_expect_tx(c_full + cst, (uint32_t)CS_BYTES);
                    umma::bulk_g2s(smemCs + (size_t)cst * CS_BYTES, prm.inv2 + ((size_t)tile * P + (size_t)j * BN), CS_BYTES, c_full + cst);
                }
                __syncwarp();
                if (++cst == CSN) { cst = 0; cph ^= 1; }
            }
        }
    } else if ((warp == 1 || (ISSUERS == 2 && warp == 2)) && rank == 0) {
        const int hsel = warp - 1;
        // ------------------------------------------------------------ MMA issuer (converged warp, elected issue; PAIR: leader CTA only)
        constexpr uint32_t idesc = umma::instr_desc_bf16(PAIR ? 2 * BM : BM, BN);
        constexpr int KPB = BK / UMMA_K;
        int bs = 0; uint32_t bph = 0, aph = 0; int acc = 0; uint32_t accph = 0;
        for (int unit = slot; unit < n_units; unit += n_slots) {
            umma::mbar_wait(a_full, aph);
            aph ^= 1;
            for (int j = 0; j < NT; ++j) {
                umma::mbar_wait(t_empty + acc, accph ^ 1);
                umma::tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    umma::mbar_wait(b_full + bs, bph);
                    umma::tc_fence_after();
                    const uint64_t bdesc = umma::smem_desc_sw128(smemB + (size_t)bs * B_STAGE_BYTES);
                    const uint64_t adesc0 = umma::smem_desc_sw128(smemA + (size_t)kb * BOX_BYTES);
                    const int nk = prm.ksteps - kb * KPB;       // padding-only K steps are skipped
#pragma unroll
                    for (int k = 0; k < KPB; ++k) {             // +32 B per K step inside the swizzle atom
                        if (k < nk) {
#pragma unroll
                            for (int h = 0; h < HALVES; ++h)
                                if ((ISSUERS == 1 || h == hsel) && umma::elect_one()) {
                                    if (PAIR) umma::mma_bf16_pair(d0 + (uint32_t)(h * 2 * BN), adesc0 + (uint64_t)(h * ((MAX_KB * BOX_BYTES) >> 4)) + 2 * k,
                                                                  bdesc + 2 * k, idesc, (kb | k) != 0);
                                    else umma::mma_bf16(d0 + (uint32_t)(h * 2 * BN), adesc0 + (uint64_t)(h * ((MAX_KB * BOX_BYTES) >> 4)) + 2 * k,
                                                        bdesc + 2 * k, idesc, (kb | k) != 0);
                                }
                        }
                    }
                    if (umma::elect_one()) { if (PAIR) umma::mma_commit_pair(b_empty + bs); else umma::mma_commit(b_empty + bs); }   // frees the B stage when these MMAs retire
                    if (++bs == NSTG) { bs = 0; bph ^= 1; }
                }
                if (umma::elect_one()) { if (PAIR) umma::mma_commit_pair(t_full + acc); else umma::mma_commit(t_full + acc); }       // accumulators of this N-tile complete
                if (++acc == 2) { acc = 0; accph ^= 1; }
            }
            if (umma::elect_one()) { if (PAIR) umma::mma_commit_pair(a_empty); else umma::mma_commit(a_empty); }                     // A may be overwritten
        }
    }
    } else {
        umma::reg_alloc<232>();
        // ------------------------------------------------------------ epilogue
        const int e = warp - FIRST_EPI_WARP;
        const int quarter = warp & 3;                   // TMEM lanes a warp may touch: 32*(warp_id % 4)..+31
        const int ch = e >> 2;                          // which half of every map row's columns
        constexpr int DH = D / 2;                       // columns of a map row handled by this thread
        constexpr int HWQ = (MODE == MODE_POOL) ? D / 4 : 1;    // pooled outputs per map row per thread
        float* stgA = smemStg + (size_t)e * (L::STG_WARP / 4);
        float* stgB = stgA + (WIDE ? 32 * 32 : 32 * STG_STRIDE);
        float4* wideA_mine = reinterpret_cast<float4*>(stgA + lane * 32);      // WIDE: rows of 32 floats, 16-byte slots XOR-swizzled by row & 7
        float4* wideB_mine = reinterpret_cast<float4*>(stgB + lane * 32);
        const int wsw = (lane >> 1) & 3;                // XOR swizzle of this lane's own staging row
        float4* stgA_mine = reinterpret_cast<float4*>(stgA + lane * STG_STRIDE);
        float4* stgB_mine = reinterpret_cast<float4*>(stgB + lane * STG_STRIDE);
        // Pair flush (MODE_POOL, D == 64 and 32).  The warps e and e + 4 own the two column halves of the same 32 + 32 patch
        // rows; flushed separately, every store instruction writes 8 rows x 64 B -- half a line per row, 8 L1
        // wavefronts for 512 bytes.  Here both warps stage into two SHARED regions instead (rows of 32 floats, the
        // 16-byte slots XOR-swizzled by row & 7: conflict-free both ways) -- the first warp's region collects
        // accumulator half A, the second warp's half B --, meet at a named barrier, and each flushes ONE region as
        // whole 128-byte lines: 4 rows x 128 B per store instruction, half the global-store wavefronts.
        // At D == 32 an N-tile holds two pooled rows of 16: the same 32 consecutive floats per patch row, a thread
        // contributing 8 + 8 of them (with LDS + STG flushes that was no faster than the per-warp flush, 3.865 vs 3.875 ms
        // on 64 x 512^2 at ws 5; it pays once the region leaves by a tensor store).
        constexpr bool PF_SHAPE = (MODE == MODE_POOL && (D == 64 || D == 32));
        const bool pf = PF_SHAPE && prm.pair_flush != 0;
        // pair_flush == 2 (the default): a staged region of 32 rows x 128 B leaves as ONE 2-D bulk tensor store -- the
        // XOR-by-row layout IS the tensor map's 128-byte swizzle -- instead of 8 LDS.128 + 8 STG.128 per lane; the
        // region is handed back (cp.async.bulk.wait_group.read) right before its next staging write.
        const bool pf_tma = PF_SHAPE && prm.pair_flush == 2;
        const bool wide_tma = WIDE && prm.pair_flush == 2;
        float* regionX = smemStg + (size_t)quarter * (2 * 32 * STG_STRIDE);
        float* regionY = smemStg + (size_t)(quarter + 4) * (2 * 32 * STG_STRIDE);
        float4* xs_mine = reinterpret_cast<float4*>(regionX + lane * 32);
        float4* ys_mine = reinterpret_cast<float4*>(regionY + lane * 32);
        int acc = 0; uint32_t accph = 0; int cst = 0; uint32_t cph = 0;
        const uint32_t t_empty_lead = PAIR ? umma::mapa_u32(t_empty, 0) : 0;     // the leader's MMA warp waits for both CTAs' epilogues
        for (int unit = slot; unit < n_units; unit += n_slots) {
            const int item = PAIR ? unit * 2 + rank : unit;
            const int tile = item / prm.items_per_tile;
            const size_t wrow = (size_t)tile * P + (size_t)(item - tile * prm.items_per_tile) * (HALVES * BM) + quarter * 32;
            const size_t prowA = wrow + lane, prowB = prowA + BM;
            const dm_stat s1A = prm.stat1[prowA], s1B = prm.stat1[prowB];
            const bool flatA = (s1A.y == 0.0f), flatB = (s1B.y == 0.0f);
            // MODE_POOL state per row: st[] = horizontally pooled previous map row (odd rows) /
            // running vertical max (even rows); rmin = running minimum of the raw values
            float stA[HWQ], stB[HWQ];
            float rminA = CUDART_INF_F, rmaxA = -CUDART_INF_F, rminB = CUDART_INF_F, rmaxB = -CUDART_INF_F;
            if (MODE == MODE_POOL) {
#pragma unroll
                for (int i = 0; i < HWQ; ++i) { stA[i] = -CUDART_INF_F; stB[i] = -CUDART_INF_F; }
            }
            // output rows of this warp are `ostride` floats apart; half B is 128 rows below half A
            const size_t ostride = (MODE == MODE_RAW) ? (size_t)P : (size_t)(P / 4);
            float* woutA = ((MODE == MODE_RAW) ? prm.raw : prm.pooled) + wrow * ostride;
            float* woutB = woutA + (size_t)BM * ostride;
            // 16 staged floats per lane -> the warp writes 8 rows x 64 B per store instruction.
            // The 16 floats are SEGS contiguous segments, `segstride` floats apart in the output.
            auto flush16 = [&](const float* stg, float* wout, size_t col, int seg, size_t segstride) {
                __syncwarp();
                const int f4 = (lane & 3) * 4;
                const uint32_t sa = umma::smem_u32(stg + (lane >> 2) * STG_STRIDE + (((lane & 3) ^ ((lane >> 3) & 3)) << 2));
                float4 v[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) v[it] = umma::lds128(sa + (uint32_t)(it * 8 * STG_STRIDE * 4));   // all four loads in flight
                float* dst = wout + (size_t)(lane >> 2) * ostride + col + (size_t)(f4 / seg) * segstride + (f4 % seg);
#pragma unroll
                for (int it = 0; it < 4; ++it) *reinterpret_cast<float4*>(dst + (size_t)(it * 8) * ostride) = v[it];
                __syncwarp();
            };
            // WIDE: 32 staged floats per lane -> the warp writes 4 rows x 128 B per store instruction
            auto flush32 = [&](const float* stg, float* wout, size_t col) {
                __syncwarp();
                const uint32_t sa = umma::smem_u32(stg);
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    float4 v[4];
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = (hb * 4 + it) * 4 + (lane >> 3);
                        v[it] = umma::lds128(sa + (uint32_t)((row * 32 + (((lane & 7) ^ (row & 7)) << 2)) * 4));
                    }
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = (hb * 4 + it) * 4 + (lane >> 3);
                        *reinterpret_cast<float4*>(wout + (size_t)row * ostride + col + (size_t)((lane & 7) << 2)) = v[it];
                    }
                }
                __syncwarp();
            };
            // pair flush: this warp writes the 32 rows x 32 floats of ITS region (ch 0: half A, ch 1: half B)
            auto flush_pair = [&](size_t col) {
                if (PF_SHAPE && pf_tma) {
                    umma::fence_proxy_async();                                        // this thread's staging writes -> async proxy
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // both column halves are staged and fenced
                    if (lane == 0) {
                        umma::tma_store_2d(&mapO, ch ? regionY : regionX, (int)col, (int)(wrow + (ch ? BM : 0)));
                        umma::bulk_commit();
                    }
                    __syncwarp();
                    return;             // the regions are released at the next N-tile's first staging write (below)
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // both column halves are staged
                const uint32_t sa = umma::smem_u32(ch ? regionY : regionX);
                float* wout = ch ? woutB : woutA;
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    float4 v[4];
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = (hb * 4 + it) * 4 + (lane >> 3);
                        v[it] = umma::lds128(sa + (uint32_t)((row * 32 + (((lane & 7) ^ (row & 7)) << 2)) * 4));
                    }
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int row = (hb * 4 + it) * 4 + (lane >> 3);
                        *reinterpret_cast<float4*>(wout + (size_t)row * ostride + col + (size_t)((lane & 7) << 2)) = v[it];
                    }
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      // the regions may be overwritten
            };
            for (int j = 0; j < NT; ++j) {
                umma::mbar_wait(c_full + cst, cph);
                umma::mbar_wait(t_full + acc, accph);
                umma::tc_fence_after();
                const uint32_t tA = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((0 * 2 + acc) * BN);
                const uint32_t tB = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((1 * 2 + acc) * BN);
                const uint32_t csm = umma::smem_u32(smemCs + (size_t)cst * CS_BYTES);
                // column of step s inside the N-tile: map row r = (8 s) / DH, then this thread's half
                auto step_col = [&](int s) -> int { return ((s * SW) / DH) * D + ch * DH + (s * SW) % DH; };
                float vA0[SW], vA1[SW], vB0[SW], vB1[SW];
                float4 c0[SW / 4], c1[SW / 4];          // inv2 of the step's 8 columns
                float hA = 0.f, hB = 0.f;               // halo column (ch == 1): raw accumulators left of the split
                float hc = 0.f;                         // its inv2
                float zprevA = -CUDART_INF_F, zprevB = -CUDART_INF_F;
                float4 obA, obB;
                {
                    const int n0 = step_col(0);
                    umma::tmem_ld_32x8_issue(tA + (uint32_t)n0, vA0);
                    umma::tmem_ld_32x8_issue(tB + (uint32_t)n0, vB0);
                    if (NORMED) {
#pragma unroll
                        for (int i = 0; i < SW / 4; ++i) c0[i] = umma::lds128(csm + 4 * n0 + 16 * i);
                    }
                    if (MODE == MODE_POOL && ch) {
                        umma::tmem_ld_32x1_issue(tA + (uint32_t)(n0 - 1), hA);
                        umma::tmem_ld_32x1_issue(tB + (uint32_t)(n0 - 1), hB);
                        if (NORMED) hc = umma::lds32(csm + 4 * (n0 - 1));
                    }
                }
#pragma unroll
                for (int s = 0; s < NSTEP; ++s) {
                    float (&vA)[SW] = (s & 1) ? vA1 : vA0;
                    float (&vB)[SW] = (s & 1) ? vB1 : vB0;
                    float (&vAn)[SW] = (s & 1) ? vA0 : vA1;
                    float (&vBn)[SW] = (s & 1) ? vB0 : vB1;
                    float4 (&cc)[SW / 4] = (s & 1) ? c1 : c0;
                    float4 (&cn)[SW / 4] = (s & 1) ? c0 : c1;
                    constexpr int dummy = 0; (void)dummy;
                    const int xo = (s * SW) % DH;           // first column of the step inside this thread's row half
                    const int r = (s * SW) / DH;            // map row inside the N-tile
                    if (WIDE && s == 0 && wide_tma && (j & 1)) {
                        if (lane == 0) umma::bulk_wait_read0();     // the previous pooled row's stores have read this warp's regions
                        __syncwarp();
                    }
                    if (PF_SHAPE && s == DH / SW && pf_tma) {       // first step of the N-tile's first odd map row
                        // the previous N-tile's tensor store has read this region (half an N-tile ago at least)
                        if (lane == 0) umma::bulk_wait_read0();
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
                    }
                    umma::tmem_wait_ld();                   // step s (and its halo) is in registers
                    // halo of THIS step must be consumed before the next step's halo load overwrites it
                    float zhA = -CUDART_INF_F, zhB = -CUDART_INF_F;
                    if (MODE == MODE_POOL && xo == 0 && ch) {
                        zhA = NORMED ? __fmul_rn(hA, hc) : hA;
                        zhB = NORMED ? __fmul_rn(hB, hc) : hB;
                    }
                    if (s + 1 < NSTEP) {                    // step s+1 in flight during the math below
                        const int n1 = step_col(s + 1);
                        umma::tmem_ld_32x8_issue(tA + (uint32_t)n1, vAn);
                        umma::tmem_ld_32x8_issue(tB + (uint32_t)n1, vBn);
                        if (NORMED) {
#pragma unroll
                            for (int i = 0; i < SW / 4; ++i) cn[i] = umma::lds128(csm + 4 * n1 + 16 * i);
                        }
                        if (MODE == MODE_POOL && ((s + 1) * SW) % DH == 0 && ch) {
                            umma::tmem_ld_32x1_issue(tA + (uint32_t)(n1 - 1), hA);
                            umma::tmem_ld_32x1_issue(tB + (uint32_t)(n1 - 1), hB);
                            if (NORMED) hc = umma::lds32(csm + 4 * (n1 - 1));
                        }
                    }
                    // the accumulators already hold dot - S1' S2'/K (correction slots): one factor per column is left
                    if (NORMED) {
#pragma unroll
                        for (int i = 0; i < SW; i += 4) {
                            const float4 cp = cc[i >> 2];       // inv2 of four columns
                            umma::mul2(vA[i], vA[i + 1], cp.x, cp.y); umma::mul2(vA[i + 2], vA[i + 3], cp.z, cp.w);
                            umma::mul2(vB[i], vB[i + 1], cp.x, cp.y); umma::mul2(vB[i + 2], vB[i + 3], cp.z, cp.w);
                        }
                    }
                    if (MODE == MODE_NULL) { rmaxA = fmaxf(rmaxA, vA[0] + vB[0]); continue; }
                    if (MODE == MODE_RAW) {
                        // D is fixed to 64 here: steps (2k, 2k+1) are 16 consecutive columns
#pragma unroll
                        for (int i = 0; i < SW; i += 4) {
                            obA.x = dm_zncc_finish(vA[i], s1A.y, flatA, NORMED); obA.y = dm_zncc_finish(vA[i + 1], s1A.y, flatA, NORMED);
                            obA.z = dm_zncc_finish(vA[i + 2], s1A.y, flatA, NORMED); obA.w = dm_zncc_finish(vA[i + 3], s1A.y, flatA, NORMED);
                            obB.x = dm_zncc_finish(vB[i], s1B.y, flatB, NORMED); obB.y = dm_zncc_finish(vB[i + 1], s1B.y, flatB, NORMED);
                            obB.z = dm_zncc_finish(vB[i + 2], s1B.y, flatB, NORMED); obB.w = dm_zncc_finish(vB[i + 3], s1B.y, flatB, NORMED);
                            stgA_mine[((s & 1) * 2 + (i >> 2)) ^ wsw] = obA;
                            stgB_mine[((s & 1) * 2 + (i >> 2)) ^ wsw] = obB;
                        }
                        if (s & 1) {
                            const size_t col = (size_t)j * BN + step_col(s - 1);
                            flush16(stgA, woutA, col, 16, 0);
                            flush16(stgB, woutB, col, 16, 0);
                        }
                    } else {
                        // map-row parity: compile time for D <= 64 (an N-tile holds BN / D rows, even);
                        // for D == 128 an N-tile is one map row and the parity alternates with j
                        const bool odd_row = (D == 128) ? ((j & 1) != 0) : ((r & 1) != 0);
#pragma unroll
                        for (int i = 0; i < SW; i += 2) {
                            const int xh = (xo + i) >> 1;   // pooled column inside this thread's half
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
                                // output index inside the group of 16 this thread completes
                                const int ol = (D == 128) ? xh : (r >> 1) * HWQ + xh;
                                if ((ol & 3) == 0) { obA.x = oA; obB.x = oB; } else if ((ol & 3) == 1) { obA.y = oA; obB.y = oB; }
                                else if ((ol & 3) == 2) { obA.z = oA; obB.z = oB; } else { obA.w = oA; obB.w = oB; }
                                if ((ol & 3) == 3) {
                                    if (PF_SHAPE && pf) {
                                        // 16-byte slot inside the row's 32 floats: D == 64: one pooled row, this half's 16 floats;
                                        // D == 32: two pooled rows of 16, this half's 8 floats of each
                                        const int k4 = (ol & 15) >> 2;
                                        const int lslot = (D == 64) ? ch * 4 + k4 : (k4 >> 1) * 4 + ch * 2 + (k4 & 1);
                                        const int slot = lslot ^ (lane & 7);
                                        xs_mine[slot] = obA; ys_mine[slot] = obB;
                                    } else if (WIDE) {
                                        wideA_mine[(ol >> 2) ^ (lane & 7)] = obA; wideB_mine[(ol >> 2) ^ (lane & 7)] = obB;
                                    } else { stgA_mine[((ol & 15) >> 2) ^ wsw] = obA; stgB_mine[((ol & 15) >> 2) ^ wsw] = obB; }
                                }
                                if (WIDE) {
                                    if (ol == 31) {         // the thread's 32 pooled floats of this pooled row: one line per patch row
                                        const size_t col = (size_t)(j >> 1) * DH + (size_t)ch * HWQ;
                                        if (wide_tma) {
                                            umma::fence_proxy_async();
                                            __syncwarp();
                                            if (lane == 0) {
                                                umma::tma_store_2d(&mapO, stgA, (int)col, (int)wrow);
                                                umma::tma_store_2d(&mapO, stgB, (int)col, (int)(wrow + BM));
                                                umma::bulk_commit();
                                            }
                                            __syncwarp();
                                        } else {
                                            flush32(stgA, woutA, col);
                                            flush32(stgB, woutB, col);
                                        }
                                    }
                                } else if ((ol & 15) == 15) {
                                    if (PF_SHAPE && pf) {
                                        flush_pair((size_t)j * (BN / 4));
                                    } else {
                                        // first pooled row / column of the group; segments of min(16, D/4) floats, one per pooled row
                                        constexpr int SEG = HWQ < 16 ? HWQ : 16;
                                        const size_t col = (D == 128) ? (size_t)(j >> 1) * DH + (size_t)ch * HWQ + (ol - 15)
                                                                      : (size_t)j * (BN / 4) + (size_t)ch * HWQ;
                                        flush16(stgA, woutA, col, SEG, DH);
                                        flush16(stgB, woutB, col, SEG, DH);
                                    }
                                }
                            }
                        }
                        zprevA = vA[SW - 1]; zprevB = vB[SW - 1];
                    }
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) umma::mbar_arrive_cluster(t_empty_lead + (uint32_t)(acc * 8)); else umma::mbar_arrive(t_empty + acc);
                    umma::mbar_arrive(c_empty + cst);
                }
                if (++acc == 2) { acc = 0; accph ^= 1; }
                if (++cst == CSN) { cst = 0; cph ^= 1; }
            }
            if (MODE == MODE_POOL) {
                // partial min / max of this column half; flat patch: OpenCV's map is all ones
                { const float v = NORMED ? (flatA ? 1.0f : fminf(fmaxf(__fmul_rn(rminA, s1A.y), -1.0f), 1.0f)) : rminA; prm.rowmin[4 * prowA + 2 * ch] = v; prm.rowmin[4 * prowA + 2 * ch + 1] = v; }
                { const float v = (NORMED && flatA) ? 1.0f : rmaxA; prm.rowmax[4 * prowA + 2 * ch] = v; prm.rowmax[4 * prowA + 2 * ch + 1] = v; }
                { const float v = NORMED ? (flatB ? 1.0f : fminf(fmaxf(__fmul_rn(rminB, s1B.y), -1.0f), 1.0f)) : rminB; prm.rowmin[4 * prowB + 2 * ch] = v; prm.rowmin[4 * prowB + 2 * ch + 1] = v; }
                { const float v = (NORMED && flatB) ? 1.0f : rmaxB; prm.rowmax[4 * prowB + 2 * ch] = v; prm.rowmax[4 * prowB + 2 * ch + 1] = v; }
            }
            if (MODE == MODE_NULL && rmaxA == 12345.678f) prm.raw[prowA] = rmaxA;     // keep the loads alive
            if ((pf_tma || wide_tma) && lane == 0) umma::bulk_wait0();                  // the unit's tensor stores are performed
        }

    }

    umma::tc_fence_before();
    __syncthreads();
    if (PAIR) umma::cluster_sync_all();     // the leader's MMAs read the peer's shared memory and signal its barriers until here
    if (warp == 1) {
        umma::tc_fence_after();
        if (PAIR) umma::tmem_dealloc_pair(tmem_base, TMEM_COLS); else umma::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

dm_encode_tiled_fn get_encode_fn() {
    static dm_encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (dm_encode_tiled_fn)p;
    return fn;
}

std::atomic<int> g_pair_mode{-1};       // -1: CTA pairs whenever the shape allows, 0: never, 1: same as -1 (dm_correlation_set_pair_mode)

template <int MODE, int D, bool NORMED, bool PAIR, bool WIDE>
int launch3(const CUtensorMap& mapA, const CUtensorMap& mapB, const Params& prm_in, cudaStream_t stream) {
    constexpr size_t SMEM_BYTES = Layout<WIDE, PAIR>::SMEM;
    Params prm = prm_in;
    CUtensorMap mapO = mapA;            // placeholder unless the pooled map leaves by tensor stores
    if (MODE == MODE_POOL && (D == 64 || D == 32 || WIDE) && prm.pair_flush == 2) {
        dm_encode_tiled_fn enc = get_encode_fn();
        DM_REQUIRE(enc != nullptr, DM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        const uint64_t rows = (uint64_t)prm.n_items * (HALVES * BM);
        cuuint64_t gdim[2] = {(cuuint64_t)(prm.P / 4), rows};
        cuuint64_t gstride[1] = {(cuuint64_t)prm.P};                 // P / 4 floats per patch row
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&mapO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, prm.pooled, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        DM_REQUIRE(r == CUDA_SUCCESS, DM_ERR_CUDA, "cuTensorMapEncodeTiled (pooled map) failed with %d", (int)r);
    }
    // per-device state: cudaFuncSetAttribute and the cluster occupancy belong to the device, and one
    // process may drive several devices from several threads (dm_multi_*).  -1 = not configured yet;
    // two threads racing on the same device compute the same values.
    constexpr int MAX_DEV = 64;
    static std::atomic<int> state[MAX_DEV];         // 0 = unconfigured, else 1 + co-resident CTA pairs
    auto kern = dm_correlation_umma_kernel<MODE, D, NORMED, PAIR, WIDE>;
    int dev = 0, sms = 0;
    DM_CUDA_CHECK(cudaGetDevice(&dev));
    DM_REQUIRE(dev >= 0 && dev < MAX_DEV, DM_ERR_UNSUPPORTED, "tcgen05 correlation: device index %d not supported", dev);
    DM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int max_pairs = state[dev].load(std::memory_order_acquire) - 1;
    if (max_pairs < 0) {
        max_pairs = 0;
        DM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        if (PAIR) {
            cudaLaunchConfig_t q = {};
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.gridDim = dim3(sms & ~1); q.blockDim = dim3(THREADS); q.dynamicSmemBytes = SMEM_BYTES; q.attrs = qa; q.numAttrs = 1;
            DM_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_pairs, kern, &q));
            DM_REQUIRE(max_pairs > 0, DM_ERR_CUDA, "tcgen05 correlation: no 2-CTA cluster fits on this device");
            if (getenv("DM_DEBUG")) fprintf(stderr, "[dm] correlation CTA pairs: %d co-resident 2-CTA clusters on %d SMs (device %d)\n", max_pairs, sms, dev);
        }
        state[dev].store(max_pairs + 1, std::memory_order_release);
    }
    if (!PAIR) {
        const int grid = prm.n_items < sms ? prm.n_items : sms;
        DM_CUDA_CHECK(dm_launch_dep(DM_PDL_CORR, kern, dim3((unsigned)grid), dim3(THREADS), SMEM_BYTES, stream, mapA, mapB, mapO, prm));
        DM_LAUNCH_CHECK();
        return DM_OK;
    }
    // persistent grid of CTA pairs: one 2-CTA cluster per TPC
    const int units = prm.n_items / 2;
    int pairs = units < max_pairs ? units : max_pairs;
    {   // measurement aid: DM_CORR_MAX_PAIRS caps the persistent grid (the other SMs stay free for a second stream)
        static const int cap = getenv("DM_CORR_MAX_PAIRS") ? atoi(getenv("DM_CORR_MAX_PAIRS")) : 0;
        if (cap > 0 && pairs > cap) pairs = cap;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;        // dm_common.cuh: the prologue overlaps the descriptor kernel's tail
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = dm_pdl_enabled(DM_PDL_CORR) ? 2 : 1;
    DM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mapA, mapB, mapO, prm));
    DM_LAUNCH_CHECK();
    return DM_OK;
}

template <int MODE, int D, bool NORMED>
int launch2(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapB_pair, const Params& prm, cudaStream_t stream) {
    // a pair's work unit is two consecutive items of the same tile
    const bool pair = g_pair_mode.load(std::memory_order_relaxed) != 0 && prm.items_per_tile % 2 == 0;
    if constexpr (MODE == MODE_POOL && D == 128) {
        static const bool no_wide = getenv("DM_CORR_NO_WIDE") != nullptr;        // measurement aid: 16-float staging rows, deeper B ring
        if (!no_wide) return pair ? launch3<MODE, D, NORMED, true, true>(mapA, mapB_pair, prm, stream)
                                  : launch3<MODE, D, NORMED, false, true>(mapA, mapB, prm, stream);
    }
    if (pair) return launch3<MODE, D, NORMED, true, false>(mapA, mapB_pair, prm, stream);
    return launch3<MODE, D, NORMED, false, false>(mapA, mapB, prm, stream);
}

template <int MODE, int D>
int launch(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapB_pair, const Params& prm, bool normed, cudaStream_t stream) {
    return normed ? launch2<MODE, D, true>(mapA, mapB, mapB_pair, prm, stream) : launch2<MODE, D, false>(mapA, mapB, mapB_pair, prm, stream);
}

}  // namespace

int dm_make_desc_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t kpad, uint32_t box_rows) {
    dm_encode_tiled_fn enc = get_encode_fn();
    DM_REQUIRE(enc != nullptr, DM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {kpad, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DM_REQUIRE(r == CUDA_SUCCESS, DM_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
    return DM_OK;
}

bool dm_correlation_umma_supported(int p, int kpad) {
    return p >= HALVES * BM && p % (HALVES * BM) == 0 && kpad % BK == 0 && kpad <= MAX_KB * BK;
}

bool dm_correlation_umma_pool_supported(int t0, int t1, int kpad) {
    return dm_correlation_umma_supported(t0 * t1, kpad) && (t1 == 16 || t1 == 32 || t1 == 64 || t1 == 128) && t0 % 2 == 0;
}

static int fill_params(Params& prm, CUtensorMap& mapA, CUtensorMap& mapB, CUtensorMap& mapB_pair, const void* desc1, const float* stat1,
                       const void* desc2, const float* stat2, int n_tiles, int p, int kpad, int kreal) {
    const uint64_t rows = (uint64_t)n_tiles * p;
    int rc = dm_make_desc_tensor_map(&mapA, desc1, rows, kpad, BM);
    if (rc != DM_OK) return rc;
    rc = dm_make_desc_tensor_map(&mapB, desc2, rows, kpad, BN);
    if (rc != DM_OK) return rc;
    rc = dm_make_desc_tensor_map(&mapB_pair, desc2, rows, kpad, BN / 2);     // each CTA of a pair loads half of the N-tile
    if (rc != DM_OK) return rc;
    prm.stat1 = (const dm_stat*)stat1;
    prm.inv2 = reinterpret_cast<const float*>((const dm_stat*)stat2 + rows);
    prm.P = p; prm.KB = kpad / BK; prm.items_per_tile = p / (HALVES * BM);
    prm.ksteps = (kreal + UMMA_K - 1) / UMMA_K;
    if (prm.ksteps <= 0 || prm.ksteps > kpad / UMMA_K) prm.ksteps = kpad / UMMA_K;
    prm.n_items = n_tiles * prm.items_per_tile;
    prm.raw = prm.pooled = prm.rowmin = prm.rowmax = nullptr;
    prm.pair_flush = 0;
    return DM_OK;
}

int dm_correlation_umma(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                        int n_tiles, int p, int kpad, int kreal, int method, float* raw, cudaStream_t stream) {
    Params prm; CUtensorMap mapA, mapB, mapBp;
    int rc = fill_params(prm, mapA, mapB, mapBp, desc1, stat1, desc2, stat2, n_tiles, p, kpad, kreal);
    if (rc != DM_OK) return rc;
    prm.raw = raw;
    return launch<MODE_RAW, 64>(mapA, mapB, mapBp, prm, method == DM_TM_CCOEFF_NORMED, stream);
}

// measurement aid (DM_CORR_UMMA_NULL): MMAs + TMEM drain, no epilogue math, no output
int dm_correlation_umma_null(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                             int n_tiles, int p, int kpad, int kreal, float* raw, cudaStream_t stream) {
    Params prm; CUtensorMap mapA, mapB, mapBp;
    int rc = fill_params(prm, mapA, mapB, mapBp, desc1, stat1, desc2, stat2, n_tiles, p, kpad, kreal);
    if (rc != DM_OK) return rc;
    prm.raw = raw;
    return launch<MODE_NULL, 64>(mapA, mapB, mapBp, prm, true, stream);
}

int dm_correlation_umma_pool(const void* desc1, const float* stat1, const void* desc2, const float* stat2,
                             int n_tiles, int t0, int t1, int kpad, int kreal, int method, int engine,
                             float* pooled, float* rowmin, float* rowmax, cudaStream_t stream) {
    DM_REQUIRE(dm_correlation_umma_pool_supported(t0, t1, kpad), DM_ERR_UNSUPPORTED, "pooled tcgen05 correlation: unsupported grid (%d,%d)", t0, t1);
    (void)engine;
    Params prm; CUtensorMap mapA, mapB, mapBp;
    int rc = fill_params(prm, mapA, mapB, mapBp, desc1, stat1, desc2, stat2, n_tiles, t0 * t1, kpad, kreal);
    if (rc != DM_OK) return rc;
    prm.pooled = pooled; prm.rowmin = rowmin; prm.rowmax = rowmax;
    static const bool no_pair_flush = getenv("DM_CORR_NO_PAIR_FLUSH") != nullptr;       // measurement aid: every warp flushes its own half rows
    static const bool no_tma_store = getenv("DM_CORR_NO_TMA_STORE") != nullptr;         // measurement aid: LDS.128 + STG.128 flushes of the staged regions
    prm.pair_flush = no_pair_flush ? 0 : (no_tma_store ? 1 : 2);
    const bool normed = method == DM_TM_CCOEFF_NORMED;
    if (t1 == 128) return launch<MODE_POOL, 128>(mapA, mapB, mapBp, prm, normed, stream);
    if (t1 == 64) return launch<MODE_POOL, 64>(mapA, mapB, mapBp, prm, normed, stream);
    if (t1 == 32) return launch<MODE_POOL, 32>(mapA, mapB, mapBp, prm, normed, stream);
    return launch<MODE_POOL, 16>(mapA, mapB, mapBp, prm, normed, stream);
}

int dm_correlation_round_tiles(int p) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = p / (HALVES * BM);
    if (items < 1) return 1;
    const bool pair = g_pair_mode.load(std::memory_order_relaxed) != 0 && items % 2 == 0;
    const int units = pair ? items / 2 : items, slots = pair ? sms / 2 : sms;
    int a = slots, b = units;
    while (b) { const int t = a % b; a = b; b = t; }       // gcd
    return slots / a;
}

void dm_correlation_umma_set_pair_mode(int mode) { g_pair_mode.store(mode, std::memory_order_relaxed); }
