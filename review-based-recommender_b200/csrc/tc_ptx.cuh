// tc_ptx.cuh — PTX wrappers shared by the tcgen05 conv kernels (conv_tc.cu: cta_group::1 / cp.async operands,
// conv_tc2.cu: cta_group::2 / TMA gather4 operands).  sm_100a only.
#pragma once
#include "rbr_common.cuh"

namespace rbr {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap), not as a hung GPU.  try_wait suspends the
// thread in hardware for a bounded time per poll, so the poll counter stays cheap (one add per failed poll).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
template <int OFF>
__device__ __forceinline__ void cp_async16_imm(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1+%2], 16, %3;" ::"r"(dst), "l"(src), "n"(OFF), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor layout):
// [0,14) start>>4 | [16,30) leading byte offset>>4 (stride between the two K core matrices) |
// [32,46) stride byte offset>>4 (stride between 8-row groups) | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=BF16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// order-preserving float → uint (and back)
__device__ __forceinline__ uint32_t f2ord(uint32_t b) { return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u); }
__device__ __forceinline__ uint32_t ord2f(uint32_t u) { return (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u; }

// Column-wise max over the 32 rows (lanes) of a warp for 16 accumulator columns held as raw fp32 bits.
// Positive floats order like signed ints, so one redux.sync.max.s32 gives the exact max whenever any candidate is >= 0;
// an all-negative column (rare: it pools to relu(...)=0 unless the bias rescues it) takes a second redux.min.u32.
// keep_v/keep_b: lane i (< 16) ends up holding column i's max bits and the ballot of rows attaining it.
template <bool FULL>
__device__ __forceinline__ void tc_colmax(const uint32_t (&v)[16], bool mine, int lane, uint32_t& keep_v, uint32_t& keep_b) {
    // the 16 reductions are independent: issued back to back so their latencies overlap (a per-column branch on the
    // all-negative case serialised them: one redux round trip per column dominated the epilogue)
    int r[16];
    int neg = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int x = (FULL || mine) ? (int)v[i] : (int)0x80000000;
        r[i] = __reduce_max_sync(0xffffffffu, x);
        neg |= r[i];
    }
    if (neg < 0) {   // warp-uniform: some column has no candidate >= 0 (common when only a few rows of the warp are `mine`)
        int rn[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t y = (FULL || mine) ? v[i] : 0xFFFFFFFFu;
            rn[i] = (int)__reduce_min_sync(0xffffffffu, y);            // smallest magnitude among negatives
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = r[i] < 0 ? rn[i] : r[i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t b = __ballot_sync(0xffffffffu, (FULL || mine) && (int)v[i] == r[i]);
        if ((lane & 15) == i) { keep_v = (uint32_t)r[i]; keep_b = b; }
    }
}



// ---- CTA-pair (cta_group::2) and TMA gather variants ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row groups 1024 bytes apart (LBO unused), layout type 2 at [61,64).
// The swizzle is a function of the absolute shared-memory address bits, so the start address may be advanced by whole
// rows (128 B) and by K-steps (32 B) inside the 1024-byte atom (probed on B200: tools/probe_umma.cu).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16_2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the MMAs issued so far retire) on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// TMA tile::gather4: 4 rows (row indices r0..r3, out-of-range → zero fill) x one box of columns starting at `col` →
// 4 consecutive 128-byte rows at `dst`; completes on the LEADER CTA's mbarrier (peer bit cleared) at offset `bar`.
__device__ __forceinline__ void tma_gather4_pair(uint32_t dst, const void* tmap, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::2 "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
        "r"(bar & 0xFEFFFFFFu)
        : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

}  // namespace rbr
