// umma_common.cuh — thin inline-PTX wrappers for the sm_100a pieces the screen kernels use: mbarrier,
// tcgen05.mma (kind::i8) with shared-memory descriptors, tcgen05.commit, tcgen05.ld, elect.sync, and their
// cta_group::2 / cluster variants.  Semantics were validated on a B200 by tools/umma_*_probe.cu.
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Bounded wait: a mis-programmed pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 26); ++it) {
        uint32_t ok;
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    __trap();
}
// one elected lane of a converged warp (the compiler keeps values used under it in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{.reg .pred P; elect.sync _|P, 0xffffffff; selp.b32 %0, 1, 0, P;}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// no-swizzle K-major shared-memory descriptor: rows linear at a 16-byte pitch (SBO = 128 B per 8 rows),
// the two 16-byte K chunks of one instruction `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)(128u >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// Wait with cluster-scope acquire: the phase may have been completed by a thread of the peer CTA whose (remote) shared-memory
// stores must be visible afterwards.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 26); ++it) {
        uint32_t ok;
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p;}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t addr_cluster, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr_cluster), "r"(v) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

// ---- CTA-pair (cta_group::2) variants ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile("{.reg .b32 ra; mapa.shared::cluster.u32 ra, %0, %1; mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];}" ::"r"(bar),
                 "r"(cta)
                 : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrive on the barrier at this offset in both CTAs
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// M = 256 across the CTA pair: each CTA supplies its own 128 rows of A and N/2 rows of B at the descriptor addresses
__device__ __forceinline__ void umma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

// The same instruction with each descriptor given as its two 32-bit halves: only the low word (start address >> 4 in bits 0-13,
// which never carries out of its field for shared-memory addresses) moves between the instructions of a tile, so an issue loop
// needs one 32-bit add per operand and instruction.
__device__ __forceinline__ void umma_i8_2cta_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{.reg .pred p; .reg .b64 da, db; setp.ne.b32 p, %6, 0; mov.b64 da, {%1, %2}; mov.b64 db, {%3, %4};"
        " tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %5, p;}" ::"r"(tmem_d),
        "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All MMAs of one window tile for KS K steps per kernel row, U kernel rows per loop iteration (KH % U == 0): KS * U instructions
// whose descriptors are independent sums off the iteration's base, so that no instruction waits for the registers of the one
// before it (a serial add -> R2UR -> UTCIMMA chain over four re-used uniform registers cost ~22 cycles per instruction).
template <int KS, int U>
__device__ __forceinline__ void umma_issue_tile(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t a_step, uint32_t b_step,
                                                uint32_t b_row, uint32_t idesc, int KH) {
    uint32_t accum = 0;
    for (int j = 0; j < KH; j += U) {
#pragma unroll
        for (int jj = 0; jj < U; ++jj) {
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                umma_i8_2cta_lohi(d, alo + (uint32_t)jj + (uint32_t)k * a_step, ahi, blo + (uint32_t)jj * b_row + (uint32_t)k * b_step, bhi, idesc, accum);
                accum = 1;
            }
        }
        alo += (uint32_t)U;
        blo += (uint32_t)U * b_row;
    }
}

// ---- TMA (cp.async.bulk.tensor) into a CTA pair's window ring ------------------------------------------------
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// one arrival + `bytes` expected transaction bytes on a barrier given by its shared::cluster address (may be the peer's)
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
// 2-D / 3-D tiled loads executed by either CTA of a pair: the box lands in THIS CTA's shared memory, the transaction bytes
// are counted on the barrier at `bar_cluster` (the leader's), as a cta_group::2 MMA needs both halves before it may issue.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void *tmap, uint32_t bar_cluster, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const void *tmap, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster)
                 : "memory");
}

// bit q of the result = any of bits 4q .. 4q+3 of m (q = 0..7): which 4-column groups of a 32-column mask are non-empty
__device__ __forceinline__ uint32_t lm_nibble_any(uint32_t m) {
    m |= m >> 1;
    m |= m >> 2;
    m &= 0x11111111u;                 // bit 4q = any of nibble q
    m = (m | (m >> 3)) & 0x03030303u; // two flags per byte
    m = (m | (m >> 6)) & 0x000f000fu; // four flags per half
    return (m | (m >> 12)) & 0xffu;
}
