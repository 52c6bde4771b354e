// PTX wrappers shared by the tcgen05 convolution kernels (conv_tc.cu, conv_tc2.cu): mbarriers, bulk async copies,
// tcgen05.mma / commit / ld, shared-memory matrix descriptors.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace aid {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// ---- CTA pair (cta_group::2) helpers: the two CTAs of a cluster issue one M = 256 MMA; the leader's commit arrives on the
// barrier at the same offset of both CTAs, the peer's warps arrive on the leader's barriers through the cluster window
__device__ __forceinline__ void tc_commit_cg2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default semantics (release at CTA scope),
// as CUTLASS' ClusterBarrier::arrive: a .release.cluster arrive makes the thread wait for its outstanding global stores to be
// performed cluster-wide, which cost the epilogue warps ~5 K cycles per tile (measured: epilogue 40-60 % slower).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B descriptor: rows of 128 bytes, 8-row groups SBO bytes apart.  The XOR pattern is a function of the
// absolute shared-memory address (bits [4,7) ^= bits [7,10)), anchored at 1024-byte boundaries, so a start address advanced
// by whole 128-byte rows or by a k offset inside the row needs no base offset (bits 49-51 stay 0).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset = 0) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}
// one elected lane of a fully active warp (the compiler then knows the guarded region runs on a single thread and can keep
// descriptors in uniform registers instead of emitting a per-lane R2UR loop around every tcgen05.mma)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}
// low / high words of a K-major SWIZZLE_128B descriptor with SBO = 1024 (see make_desc_sw128)
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
static constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);

// K-major SWIZZLE_64B (rows of 64 bytes, 8-row groups 512 bytes apart, XOR of address bits [4,6) with bits [7,9)): high word
static constexpr uint32_t DESC_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);

// tcgen05.mma with the descriptors given as their low words (the high word of a K-major SWIZZLE_128B descriptor with SBO = 1024 is
// a constant): the 64-bit descriptors are assembled inside the asm, so a uniform low word stays in a uniform register
__device__ __forceinline__ void tc_mma_f16_lo(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(DESC_HI_SW128)
        : "memory");
}
// NK (4 or 2) consecutive 16-channel k-steps of a 64-channel group (descriptor start advanced by 32 bytes = 2 units each) as
// one asm block, for one accumulator or for two accumulators that share the B operand (the two pixel units of a tile,
// interleaved unit 0 / unit 1).  CG = 1: cta_group::1 (M = 128); CG = 2: cta_group::2 (M = 256 over a CTA pair, issued by
// the leader, the descriptors address the same shared-memory offsets in both CTAs).
#define AID_MMA1(CG, P) "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %3, " P ";\n\t"
#define AID_MMA1_STEP(CG, OFF) "add.u32 a, %1, " OFF ";\n\t add.u32 b, %2, " OFF ";\n\t mov.b64 da, {a, %5};\n\t mov.b64 db, {b, %5};\n\t" AID_MMA1(CG, "q")
#define AID_MMA1_HEAD(CG)                                                                                            \
    "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"                                              \
    "setp.ne.b32 p, %4, 0;\n\t setp.eq.b32 q, 0, 0;\n\t mov.b64 da, {%1, %5};\n\t mov.b64 db, {%2, %5};\n\t" AID_MMA1(CG, "p")
template <int CG, int NK>
__device__ __forceinline__ void tc_mma_k(uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
    static_assert((CG == 1 || CG == 2) && (NK == 2 || NK == 4), "tc_mma_k");
#define AID_OPS1 ::"r"(d), "r"(alo), "r"(blo), "r"(idesc), "r"(acc), "r"(DESC_HI_SW128) : "memory"
    if constexpr (CG == 1 && NK == 4) asm volatile(AID_MMA1_HEAD("1") AID_MMA1_STEP("1", "2") AID_MMA1_STEP("1", "4") AID_MMA1_STEP("1", "6") "}" AID_OPS1);
    if constexpr (CG == 1 && NK == 2) asm volatile(AID_MMA1_HEAD("1") AID_MMA1_STEP("1", "2") "}" AID_OPS1);
    if constexpr (CG == 2 && NK == 4) asm volatile(AID_MMA1_HEAD("2") AID_MMA1_STEP("2", "2") AID_MMA1_STEP("2", "4") AID_MMA1_STEP("2", "6") "}" AID_OPS1);
    if constexpr (CG == 2 && NK == 2) asm volatile(AID_MMA1_HEAD("2") AID_MMA1_STEP("2", "2") "}" AID_OPS1);
#undef AID_OPS1
}
#define AID_MMA2(CG, P0, P1) "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da0, db, %5, " P0 ";\n\t" "tcgen05.mma.cta_group::" CG ".kind::f16 [%1], da1, db, %5, " P1 ";\n\t"
#define AID_MMA2_STEP(CG, OFF)                                                                                       \
    "add.u32 a0, %2, " OFF ";\n\t add.u32 a1, %3, " OFF ";\n\t add.u32 b, %4, " OFF ";\n\t"                          \
    "mov.b64 da0, {a0, %8};\n\t mov.b64 da1, {a1, %8};\n\t mov.b64 db, {b, %8};\n\t" AID_MMA2(CG, "q", "q")
#define AID_MMA2_HEAD(CG)                                                                                            \
    "{\n\t.reg .pred p0, p1, q;\n\t.reg .b64 da0, da1, db;\n\t.reg .b32 a0, a1, b;\n\t"                              \
    "setp.ne.b32 p0, %6, 0;\n\t setp.ne.b32 p1, %7, 0;\n\t setp.eq.b32 q, 0, 0;\n\t"                                 \
    "mov.b64 da0, {%2, %8};\n\t mov.b64 da1, {%3, %8};\n\t mov.b64 db, {%4, %8};\n\t" AID_MMA2(CG, "p0", "p1")
template <int CG, int NK>
__device__ __forceinline__ void tc_mma_kx2(uint32_t d0, uint32_t d1, uint32_t alo0, uint32_t alo1, uint32_t blo, uint32_t idesc, uint32_t acc0,
                                           uint32_t acc1) {
    static_assert((CG == 1 || CG == 2) && (NK == 2 || NK == 4), "tc_mma_kx2");
#define AID_OPS2 ::"r"(d0), "r"(d1), "r"(alo0), "r"(alo1), "r"(blo), "r"(idesc), "r"(acc0), "r"(acc1), "r"(DESC_HI_SW128) : "memory"
    if constexpr (CG == 1 && NK == 4) asm volatile(AID_MMA2_HEAD("1") AID_MMA2_STEP("1", "2") AID_MMA2_STEP("1", "4") AID_MMA2_STEP("1", "6") "}" AID_OPS2);
    if constexpr (CG == 1 && NK == 2) asm volatile(AID_MMA2_HEAD("1") AID_MMA2_STEP("1", "2") "}" AID_OPS2);
    if constexpr (CG == 2 && NK == 4) asm volatile(AID_MMA2_HEAD("2") AID_MMA2_STEP("2", "2") AID_MMA2_STEP("2", "4") AID_MMA2_STEP("2", "6") "}" AID_OPS2);
    if constexpr (CG == 2 && NK == 2) asm volatile(AID_MMA2_HEAD("2") AID_MMA2_STEP("2", "2") "}" AID_OPS2);
#undef AID_OPS2
}
// two k-steps of a 32-channel group stored with 64-byte rows (SWIZZLE_64B descriptors), cta_group::1
__device__ __forceinline__ void tc_mma_k2_sw64(uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t.reg .b32 a, b;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t setp.eq.b32 q, 0, 0;\n\t mov.b64 da, {%1, %5};\n\t mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.u32 a, %1, 2;\n\t add.u32 b, %2, 2;\n\t mov.b64 da, {a, %5};\n\t mov.b64 db, {b, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, q;\n\t}" ::"r"(d), "r"(alo), "r"(blo), "r"(idesc), "r"(acc), "r"(DESC_HI_SW64)
        : "memory");
}
// one (kf, group) stage of the 5x3 layers: the three kt taps (activation descriptor advanced by one 128-byte pixel row = 8
// units, weight descriptor by ktd) for the accumulators whose unit is in range (v0 / v1)
template <int CG, int NK>
__device__ __forceinline__ void tc_stage3(uint32_t d0, uint32_t d1, uint32_t a0, uint32_t a1, uint32_t b, uint32_t ktd, uint32_t idesc, bool v0,
                                          bool v1, uint32_t acc0, uint32_t acc1) {
    if (v0 && v1) {
        tc_mma_kx2<CG, NK>(d0, d1, a0, a1, b, idesc, acc0, acc1);
        tc_mma_kx2<CG, NK>(d0, d1, a0 + 8u, a1 + 8u, b + ktd, idesc, 1u, 1u);
        tc_mma_kx2<CG, NK>(d0, d1, a0 + 16u, a1 + 16u, b + 2u * ktd, idesc, 1u, 1u);
    } else if (v0) {
        tc_mma_k<CG, NK>(d0, a0, b, idesc, acc0);
        tc_mma_k<CG, NK>(d0, a0 + 8u, b + ktd, idesc, 1u);
        tc_mma_k<CG, NK>(d0, a0 + 16u, b + 2u * ktd, idesc, 1u);
    } else if (v1) {
        tc_mma_k<CG, NK>(d1, a1, b, idesc, acc1);
        tc_mma_k<CG, NK>(d1, a1 + 8u, b + ktd, idesc, 1u);
        tc_mma_k<CG, NK>(d1, a1 + 16u, b + 2u * ktd, idesc, 1u);
    }
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4p_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8p_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

}  // namespace aid
