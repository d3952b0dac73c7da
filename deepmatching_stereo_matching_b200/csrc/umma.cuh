// Thin inline-PTX layer for Blackwell (sm_100a): mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05.mma / TMEM allocation / tcgen05.ld, and the UMMA shared-memory / instruction
// descriptors.  Hand-written; nothing here comes from a library.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp.  The single-thread instructions (tcgen05.mma, tcgen05.commit,
// TMA) are issued as `if (elect_one()) ...` from warp-uniform control flow: ptxas then keeps
// addresses and descriptors in uniform registers and emits the UTCHMMA / UTMALDG back to back.
// Behind a per-lane branch (`if (lane == 0)`) every one of them is wrapped in an
// ELECT / BRA.U.ANY loop, which made the MMA issue rate the bottleneck at N = 64.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tile load: c0 = innermost (element) coordinate, c1 = row coordinate
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 2-D tile store shared -> global (bulk async-group completion): c0 = innermost coordinate, c1 = row coordinate
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // writes performed
// 1-D bulk copy global -> shared
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- register re-partitioning
// (all threads of a warpgroup must execute these; counts are multiples of 8)
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {       // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {                                     // whole warp
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {          // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tcgen05.ld 32x32b: thread i of the warp receives N consecutive columns of TMEM lane
// (lane_base + i).  A warp may only touch lanes 32*(warp_id % 4) .. +31.  The load is
// asynchronous: issue it, overlap independent work, then tmem_wait_ld() before using v.
// NOTE: a TMEM accumulator stage must only be handed back to the MMA warp (mbarrier arrive)
// after every value loaded from it has been consumed -- releasing right after the wait let
// the next MMA overwrite the stage under the tail of an in-flight load (see
// tests/test_gpu_parity.py::test_correlation_engines_are_repeatable).
__device__ __forceinline__ void tmem_ld_32x8_issue(uint32_t taddr, float (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1_issue(uint32_t taddr, float& v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v) : "r"(taddr) : "memory");
}

// 16-byte shared load pinned in program order (volatile): keeps the software pipeline's
// prefetches where they were written instead of letting the scheduler sink them to the use
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}

// ---------------------------------------------------------------- packed / 3-input fp32 (sm_100)
// (v0,v1) = (fma(ns1, s2k, v) * inv) for two columns at once: FFMA2 + FMUL2, each lane IEEE-rn,
// bit-identical to dm_zncc_partial.
__device__ __forceinline__ void zncc_partial2(float& v0, float& v1, float ns1, float s2k0, float s2k1, float inv0, float inv1) {
    asm("{ .reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 rb, {%5, %6};\n\t"
        "mul.rn.f32x2 rc, rc, rb;\n\t"
        "mov.b64 {%0, %1}, rc; }"
        : "+f"(v0), "+f"(v1) : "f"(ns1), "f"(s2k0), "f"(s2k1), "f"(inv0), "f"(inv1));
}
// (v0, v1) *= (f0, f1): FMUL2, each lane IEEE-rn
__device__ __forceinline__ void mul2(float& v0, float& v1, float f0, float f1) {
    asm("{ .reg .b64 ra, rb;\n\t"
        "mov.b64 ra, {%0, %1};\n\t"
        "mov.b64 rb, {%2, %3};\n\t"
        "mul.rn.f32x2 ra, ra, rb;\n\t"
        "mov.b64 {%0, %1}, ra; }"
        : "+f"(v0), "+f"(v1) : "f"(f0), "f"(f1));
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---------------------------------------------------------------- MMA
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 bf16
// (128 B), 8-row groups 1024 B apart (SBO), LBO unused (=1), descriptor version 1.
__device__ __forceinline__ uint64_t smem_desc_sw128(const void* tile) {
    const uint32_t addr = smem_u32(tile);
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                  // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                  // version (Blackwell)
    d |= (uint64_t)2 << 61;                  // layout type SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: BF16 x BF16 -> F32, both operands K-major
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by one thread for the whole CTA
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// mbarrier arrive once all MMAs issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ---------------------------------------------------------------- CTA pair (cta_group::2)
// Two CTAs of one cluster (one TPC) execute ONE tcgen05.mma together: M = 256 (128 TMEM lanes
// in each CTA), each CTA supplies its own 128 rows of A and HALF of the B tile from its own
// shared memory, so the operand fetch per CTA drops from A + B to A + B/2.  Only the leader
// (cluster rank 0) issues MMAs and commits; barriers that the leader waits on live in the
// leader's shared memory and are signalled remotely by the peer.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {           // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (an object in this CTA's shared memory) inside CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// Remote arrive with the default .release.cta semantics.  (.release.cluster costs a cluster-scope
// MEMBAR + L1 invalidate per arrive: with it the epilogue warps spent a third of their time waiting
// for their own global stores to drain.  The TMEM reads it orders are already complete after
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA tile load into THIS CTA's shared memory, completion bytes reported to a barrier that may
// live in the peer CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // one warp (same index) in BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA] * B[N: N/2 rows from each CTA]^T; leader only
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far retired
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

}  // namespace umma

// ---------------------------------------------------------------- host: tensor maps
typedef CUresult (*dm_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows][kpad] bf16 row-major descriptor matrix, box = 64 elements (128 B) x box_rows,
// 128-byte swizzle.  Returns 0 on success.
int dm_make_desc_tensor_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t kpad, uint32_t box_rows);
