// tcgen05 implicit-GEMM convolution, second generation (conv_mode 2): single-fp16 operands, channels-last operand
// layout, large bulk copies, separate activation / weight rings, 32-column epilogue batches.
//
// Same GEMM view, unit definition and fused epilogue as conv_tc.cu (out = alpha*(acc*gate + R), group statistics):
//     D[128 px, Ntile couts] += A[128 px, 16 ch] * B[Ntile, 16 ch]^T     for every (kf, kt, 16-channel step)
// What changed, and why (measured with tools/bench_bulk.cu on B200: every cp.async.bulk costs ~75-150 cycles of
// issue/processing regardless of its size, every mbarrier operation ~100 cycles, L2 -> shared tops out near 70 B/clk/SM):
//   * activations in HBM : [B][G = ceil(C/64)][rows][T+2][64] fp16, 128-byte pixel rows whose eight 16-byte chunks are
//     stored XOR-swizzled by (flattened padded pixel index & 7).  A unit's 130-pixel window of one 64-channel group is ONE
//     contiguous 16.6 KB run: a single bulk copy lands it in shared memory, already in the K-major SWIZZLE_128B UMMA
//     layout (the copy starts at row (pixel index & 7) of a 1024-byte aligned slot so that shared-memory row phase ==
//     global pixel phase).  The kt taps are the same window with the descriptor start advanced by one 128-byte row, the
//     four 16-channel k-steps of a group advance it by 32 bytes.  (conv_tc.cu needs eight 2 KB copies for the same data.)
//   * weights in HBM     : [n-tile][kf][G][kt][Ntile][64] fp16 (x 2^10), chunks swizzled by (n & 7): one copy per
//     (kf, group, kt-chunk).
//   * couts per tile: min(Cout, 256), except multi-tap layers with 256 couts, which run as two 128-wide n-tiles so that the
//     accumulators stay double buffered (tc2_ntile); tiles are numbered n-tile-minor then.
//   * two producer warps (activations, weights) with their own rings, so a stage costs one wait + one expect + 1-2 copies;
//   * epilogue: 8 warps, each thread owns one pixel and walks its warp's column range 32 columns at a time with all 32
//     residual loads in flight before the accumulator is read (the old 8-column batches left the LSU latency bound).
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace aid {

// warps 0 .. EW-1: epilogue, EW: activation producer, EW+1: weight producer, EW+2: cta_group::2 relay (else idle), EW+3: MMA issuer.
// The SM's warp arbiter favours the highest warp id of a sub-partition, and the MMA issuer is the one serial, latency-critical
// instruction stream.  EW = 16 (four column groups x four TMEM lane quadrants, 16-column batches, <= 96 registers per thread)
// when the n-tile splits into four column groups of whole 8-column chunks, else EW = 8 (two column groups, 32-column batches).
// The epilogue is a dependent instruction stream per thread (address, load, FMA, store, statistics): what it needs is more
// warps to interleave, not more bytes in flight per warp.
static constexpr int T2_EPI_WARP0 = 0;
static constexpr int T2_BAR_BYTES = 512;        // mbarriers + TMEM slot at the end of the rings
static constexpr int T2_STAT_GATE_BYTES = 4096;  // the epilogue warps' private gate tables (EW x 256 / NCW floats)
static constexpr int T2_STAT_SMEM = 16384;       // their double statistics accumulators
static constexpr int T2_ASLOT_UNIT = 18 * 1024; // one unit's window: (130 + 7) rows x 128 B, rounded to 1 KB

struct Tc2Args {
    const __half* a; const __half* w;
    TV out, R;
    const float* gate; long long gate_bstride;
    float alpha; double* stats;
    int B, Cin, G, Ntot, Ntile, n_ntiles, F, T, Tp, dil;
    int PF, rows_total, stream, units_per_b;
    int KF, KT, kt_shift, ktb;                 // ktb: kt taps per weight slot
    int tiles_t, n_units, n_pairs, n_tiles;
    int nA, nB, b_slot_bytes, acc_bufs, ncol_stride;
    uint32_t mg_pairs, mg_upb, mg_Tp, mg_tt, mg_F;   // division magics (fast_divmod) of n_pairs, units_per_b, Tp, tiles_t, F
    int cg2;            // 1: cta_group::2 -- a cluster of two CTAs works on a quad of four units (two each) with M = 256 MMAs issued by the
                        //    leader; each CTA holds half of the couts of a weight slot
    int qmode;          // cg2: how the four units of a quad are chosen (unit_index)
    int a_slot_bytes;   // activation ring slot: 2 unit windows
    int epi_class;      // 0: generic epilogue, 1-5: epilogue_fast<BW, NB, GCN> instantiation (launch_conv_tc2)
    int nt_minor;       // 1: tile = 2 * pair + n-tile (two n-tiles), else tile = n-tile * n_pairs + pair
    int out_cl, r_cl;   // 1: that tensor is channels-last [B][F][T][C] (C = its TV's channel count), else NCHW
    int dbg;  // AID_TC_DEBUG bits (tuning only): 1 skip epilogue body, 2 skip MMAs, 4 skip A loads, 8 skip B loads
};

struct Unit2 { int exists, b, f_lo, f_hi, win_start, o0; };

// Pipeline profile (AID_TC_DEBUG bit 2048, tuning only): cycles each role spent waiting / working, summed over the CTAs.
//   0 MMA wait tmem_empty, 1 MMA wait a_full (cta_group::2: wait ready), 2 MMA wait b_full, 3 MMA issue + commit, 4 MMA total,
//   5 A producer wait a_empty, 6 B producer wait b_empty, 7 epilogue warp 0 wait tmem_full, 8 epilogue warp 0 total, 9 CTAs
__device__ unsigned long long g_tc2_prof[16];
#define T2_PROF_T0(var) long long var = 0; if (prof) var = clock64()
#define T2_PROF_ADD(slot, var) do { if (prof) prof_acc[slot] += clock64() - var; } while (0)

// one unit window of zeros (130 pixels x 128 B): what a CTA of a pair loads when a tap row is outside its own unit's plane but
// inside its partner's (the M = 256 MMA covers both units)
__device__ uint4 g_tc2_zero_window[130 * 8];

// n / d and n % d with a host-computed magic = floor(2^32 / d) (0xffffffff for d == 1): the estimate is low by at most one
__device__ __forceinline__ uint2 fast_divmod(uint32_t n, uint32_t d, uint32_t magic) {
    uint32_t q = __umulhi(n, magic), r = n - q * d;
    if (r >= d) { ++q; r -= d; }
    return make_uint2(q, r);
}
static uint32_t div_magic(uint32_t d) { return d <= 1 ? 0xffffffffu : (uint32_t)(0x100000000ull / d); }

__device__ __forceinline__ Unit2 unit2_info(const Tc2Args& p, int u) {
    Unit2 i;
    i.exists = u < p.n_units;
    if (p.stream) {
        const uint2 bk = fast_divmod((uint32_t)u, (uint32_t)p.units_per_b, p.mg_upb);
        i.b = (int)bk.x;
        i.o0 = (int)bk.y * 128;                       // first output position in the padded stream of the real rows
        i.f_lo = (int)fast_divmod((uint32_t)i.o0, (uint32_t)p.Tp, p.mg_Tp).x;
        i.f_hi = min(p.F - 1, (int)fast_divmod((uint32_t)(i.o0 + 127), (uint32_t)p.Tp, p.mg_Tp).x);
        i.win_start = p.PF * p.Tp + i.o0 - 1;         // window = positions [o0-1, o0+129) of the padded plane
    } else {
        const uint2 rt = fast_divmod((uint32_t)u, (uint32_t)p.tiles_t, p.mg_tt);
        const uint2 bf = fast_divmod(rt.x, (uint32_t)p.F, p.mg_F);
        const int f = (int)bf.y, t0 = (int)rt.y * 128;
        i.b = (int)bf.x; i.f_lo = i.f_hi = f;
        i.o0 = f * p.Tp + t0 + 1;
        i.win_start = f * p.Tp + t0;
    }
    return i;
}

// tile -> (n-tile, unit pair / quad).  nt_minor (two n-tiles): neighbouring CTAs (clusters) work on the two n-tiles of the same
// units at the same time, so the second read of the activation windows hits L2, and with an even grid a CTA keeps one half of
// the weights.
__device__ __forceinline__ uint2 tile_decode(const Tc2Args& p, int tile) {
    if (p.nt_minor) return make_uint2((uint32_t)tile & 1u, (uint32_t)tile >> 1);
    return fast_divmod((uint32_t)tile, (uint32_t)p.n_pairs, p.mg_pairs);
}

// Unit j (0, 1) of CTA r.  cta_group::1: r = 0, tile index = unit pair.  cta_group::2: the tile is a quad of four units, two per CTA
// of the pair; unit j of both CTAs feeds one M = 256 MMA through ONE descriptor, so the two windows must sit at the same
// shared-memory row phase (win_start & 7):
//   qmode 0: four consecutive units (same row when tiles_t % 4 == 0; stream mode: every unit has the same phase)
//   qmode 1: tiles_t == 2: the two t-tiles of row f (leader) and of row f + 4 (peer) of an 8-row block (4 * Tp = 0 mod 8)
//   qmode 2: tiles_t == 1: rows 2h + j (leader) and 2h + j + 4 (peer) of an 8-row block, h = quad & 1
__device__ __forceinline__ int unit_index(const Tc2Args& p, int tp, int r, int j) {
    if (!p.cg2) return 2 * tp + j;
    if (p.qmode == 0) return 4 * tp + 2 * r + j;
    if (p.qmode == 1) return (((tp >> 2) * 8 + (tp & 3) + 4 * r) << 1) + j;
    return (tp >> 1) * 8 + 2 * (tp & 1) + j + 4 * r;
}

// per-kf validity bits of a tile: own0 / own1 = this CTA's units, any0 / any1 = this CTA's or (cta_group::2) the partner's unit j.
// A (kf, group) stage exists when (any0 | any1) has bit kf; every role of both CTAs walks the same stage sequence.
struct TileUnits { Unit2 u0, u1; uint32_t own0, own1, any0, any1; };
__device__ __forceinline__ uint32_t rows_kf_mask(const Tc2Args& p, int f_lo, int f_hi) {
    uint32_t m = 0;
    for (int kf = 0; kf < p.KF; ++kf) {
        const int foff = (kf - p.KF / 2) * p.dil;
        m |= (f_hi + foff >= 0 && f_lo + foff < p.F ? 1u : 0u) << kf;
    }
    return m;
}
__device__ __forceinline__ TileUnits tile_units(const Tc2Args& p, int tp, int r) {
    TileUnits t;
    t.u0 = unit2_info(p, unit_index(p, tp, r, 0));
    t.u1 = unit2_info(p, unit_index(p, tp, r, 1));
    t.own0 = t.u0.exists ? rows_kf_mask(p, t.u0.f_lo, t.u0.f_hi) : 0u;
    t.own1 = !t.u1.exists ? 0u : ((!p.stream && t.u1.f_lo == t.u0.f_lo) ? t.own0 : rows_kf_mask(p, t.u1.f_lo, t.u1.f_hi));
    t.any0 = t.own0; t.any1 = t.own1;
    if (p.cg2) {
        if (p.stream) {
            const Unit2 q0 = unit2_info(p, unit_index(p, tp, r ^ 1, 0)), q1 = unit2_info(p, unit_index(p, tp, r ^ 1, 1));
            if (q0.exists) t.any0 |= rows_kf_mask(p, q0.f_lo, q0.f_hi);
            if (q1.exists) t.any1 |= rows_kf_mask(p, q1.f_lo, q1.f_hi);
        } else if (p.qmode != 0) {      // the partner's units are 4 rows below (leader) / above (peer) ours
            const int df = r ? -4 : 4;
            t.any0 |= rows_kf_mask(p, t.u0.f_lo + df, t.u0.f_lo + df);
            t.any1 |= rows_kf_mask(p, t.u1.f_lo + df, t.u1.f_lo + df);
        }                               // qmode 0: the four units share a row
    }
    return t;
}

// unit sequence of one epilogue thread of conv_tc2 (tile -> its one or two units), for epilogue_fast
struct Tc2UnitIter {
    const Tc2Args& p; int crank, tstep, pofs;
    int it_tile, it_ui = 0, it_ab = 0; uint32_t it_aph = 0;
    __device__ Tc2UnitIter(const Tc2Args& p_, int crank_, int tile0, int tstep_, int pofs_) : p(p_), crank(crank_), tstep(tstep_), pofs(pofs_), it_tile(tile0) {}
    __device__ __forceinline__ bool next(EpiUnit& d) {
        if (it_tile >= p.n_tiles) return false;
        uint32_t pair = (uint32_t)it_tile; d.nt = 0;
        if (p.n_ntiles > 1) { const uint2 dm = tile_decode(p, it_tile); d.nt = (int)dm.x; pair = dm.y; }
        const bool has1 = unit_index(p, (int)pair, crank, 1) < p.n_units;
        Unit2 u = unit2_info(p, unit_index(p, (int)pair, crank, it_ui));
        if (!u.exists) u.b = 0;                   // cta_group::2, last quad: this CTA only keeps the handshakes going
        const int o = u.o0 + pofs;                // output position in the padded stream of the real rows
        const int row = (int)fast_divmod((uint32_t)o, (uint32_t)p.Tp, p.mg_Tp).x, tp = o - row * p.Tp;
        d.ok = u.exists && tp >= 1 && tp <= p.T && row <= u.f_hi;
        // lanes that own no real pixel read (never write) pixel 0 of their clip: the loads need no predicate
        d.pix = d.ok ? (long long)row * p.T + (tp - 1) : 0;
        d.b = u.b; d.ab = it_ab; d.aph = it_aph;
        d.tcol = (uint32_t)(it_ab * 2 * p.ncol_stride + it_ui * p.ncol_stride);
        d.first = it_ui == 0;
        d.last = it_ui == 1 || !has1;
        if (d.last) {
            it_ui = 0; it_tile += tstep;
            if (++it_ab == p.acc_bufs) { it_ab = 0; it_aph ^= 1; }
        } else it_ui = 1;
        return true;
    }
};

// CG2 = false: cta_group::1, plain launch.  CG2 = true: cta_group::2, launched in clusters of two CTAs (a kernel that contains
// cta_group::2 instructions cannot be launched without a cluster, hence two instantiations).
template <bool CG2, int EW>
__device__ __forceinline__ void conv_tc2_body(const Tc2Args& p) {
    constexpr int T2_EPI_WARPS = EW, T2_WARP_A = EW, T2_WARP_B = EW + 1, T2_WARP_RELAY = EW + 2, T2_WARP_MMA = EW + 3;
    constexpr int NCW = EW / 4;                 // column groups of an n-tile (one per 4 epilogue warps)
    constexpr int BW = EW == 8 ? 32 : 16;       // columns per epilogue batch
    constexpr int GP = 8 / NCW;                 // statistics groups a warp's columns can span
    constexpr int GSM_W = 256 / NCW;            // floats of a warp's private gate table
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1 KB aligned, still a shared-space pointer for the compiler
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_slot_bytes = p.a_slot_bytes;
    constexpr bool cg2 = CG2;
    uint32_t crank = 0;
    if constexpr (cg2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    // cta_group::2: a tile belongs to the cluster; tiles advance by cluster
    const int tile0 = cg2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tstep = cg2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    uint8_t* ringA = smem;
    uint8_t* ringB = smem + (size_t)p.nA * a_slot_bytes;
    uint8_t* bar_base = ringB + (size_t)p.nB * p.b_slot_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* a_empty = a_full + 4;
    uint64_t* b_full = a_empty + 4;
    uint64_t* b_empty = b_full + 8;
    uint64_t* tmem_full = b_empty + 8;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* ready = tmem_empty + 2;          // cta_group::2, leader: both CTAs' operands of weight slot s have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + 8);

    if (warp == T2_WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < p.nA; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
            for (int s = 0; s < p.nB; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); mbar_init(ready + s, 2); }
            for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, cg2 ? 2 * EW : EW); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if constexpr (cg2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if constexpr (cg2) cluster_sync_all();   // the partner's barriers must be initialised before commits / arrivals can reach them
    const uint32_t tmem_base = *tmem_slot;
    const int nktb = (p.KT + p.ktb - 1) / p.ktb;   // weight slots per (kf, group)

    if (warp == T2_WARP_A) {
        // ===================== activation producer: one 16.6 KB bulk copy per (unit, kf, 64-channel group) =====================
        int slot = 0; uint32_t phase = 0;
        const bool prof = (p.dbg & 2048) != 0;
        long long prof_acc[1] = {0};
        const size_t gstride = (size_t)p.rows_total * p.Tp * 64;   // halves per (clip, group) plane
        for (int tile = tile0; tile < p.n_tiles; tile += tstep) {
            const TileUnits tu = tile_units(p, (int)tile_decode(p, tile).y, (int)crank);
            for (int kf = 0; kf < p.KF; ++kf) {
                if (!(((tu.any0 | tu.any1) >> kf) & 1u)) continue;
                const int foff = (kf - p.KF / 2) * p.dil;
                // c: something lands in the unit's half of the slot; v: the unit's own window (else a window of zeros)
                const bool c0 = (tu.any0 >> kf) & 1u, c1 = (tu.any1 >> kf) & 1u, v0 = (tu.own0 >> kf) & 1u, v1 = (tu.own1 >> kf) & 1u;
                const int s0 = tu.u0.win_start + foff * p.Tp, s1 = tu.u1.win_start + foff * p.Tp;
                for (int g = 0; g < p.G; ++g) {
                    T2_PROF_T0(t_w);
                    mbar_wait(a_empty + slot, phase ^ 1);
                    T2_PROF_ADD(0, t_w);
                    if (lane == 0) {
                        uint8_t* sa = ringA + (size_t)slot * a_slot_bytes;
                        const uint32_t bytes = (p.dbg & 4) ? 0u : ((c0 ? 130u * 128u : 0u) + (c1 ? 130u * 128u : 0u));
                        mbar_expect_tx(a_full + slot, bytes);
                        if (!(p.dbg & 4)) {
                            if (c0) bulk_g2s(sa + (s0 & 7) * 128, v0 ? (const void*)(p.a + ((size_t)tu.u0.b * p.G + g) * gstride + (size_t)s0 * 64) : (const void*)g_tc2_zero_window,
                                             130u * 128u, a_full + slot);
                            if (c1) bulk_g2s(sa + T2_ASLOT_UNIT + (s1 & 7) * 128, v1 ? (const void*)(p.a + ((size_t)tu.u1.b * p.G + g) * gstride + (size_t)s1 * 64) : (const void*)g_tc2_zero_window,
                                             130u * 128u, a_full + slot);
                        }
                    }
                    __syncwarp();
                    if (++slot == p.nA) { slot = 0; phase ^= 1; }
                }
            }
        }
        if (prof && lane == 0) atomicAdd(&g_tc2_prof[5], (unsigned long long)prof_acc[0]);
    } else if (warp == T2_WARP_B) {
        // ===================== weight producer: one bulk copy per (kf, group, kt chunk) =====================
        // cta_group::2: this CTA holds couts [crank * Ntile / 2, +Ntile / 2) of every kt sub-tile (three copies per slot)
        int slot = 0; uint32_t phase = 0;
        const bool prof = (p.dbg & 2048) != 0;
        long long prof_acc[1] = {0};
        const size_t kt_halves = (size_t)p.Ntile * 64;
        for (int tile = tile0; tile < p.n_tiles; tile += tstep) {
            const uint2 tdm = tile_decode(p, tile);
            const int nt = (int)tdm.x;
            const TileUnits tu = tile_units(p, (int)tdm.y, (int)crank);
            for (int kf = 0; kf < p.KF; ++kf) {
                if (!(((tu.any0 | tu.any1) >> kf) & 1u)) continue;
                for (int g = 0; g < p.G; ++g) {
                    const __half* wg = p.w + (((size_t)nt * p.KF + kf) * p.G + g) * p.KT * kt_halves;
                    for (int c = 0; c < nktb; ++c) {
                        const int nkt = min(p.ktb, p.KT - c * p.ktb);
                        T2_PROF_T0(t_w);
                        mbar_wait(b_empty + slot, phase ^ 1);
                        T2_PROF_ADD(0, t_w);
                        if (lane == 0) {
                            if (!cg2) {
                                const uint32_t bytes = (uint32_t)(nkt * kt_halves * 2);
                                mbar_expect_tx(b_full + slot, (p.dbg & 8) ? 0u : bytes);
                                if (!(p.dbg & 8)) bulk_g2s(ringB + (size_t)slot * p.b_slot_bytes, wg + (size_t)c * p.ktb * kt_halves, bytes, b_full + slot);
                            } else {
                                const uint32_t half = (uint32_t)kt_halves;      // bytes of half a kt sub-tile (Ntile / 2 rows of 128 B)
                                mbar_expect_tx(b_full + slot, (p.dbg & 8) ? 0u : (uint32_t)nkt * half);
                                if (!(p.dbg & 8))
                                    for (int k = 0; k < nkt; ++k)
                                        bulk_g2s(ringB + (size_t)slot * p.b_slot_bytes + (size_t)k * half,
                                                 reinterpret_cast<const uint8_t*>(wg + (size_t)k * kt_halves) + crank * half, half, b_full + slot);
                            }
                        }
                        __syncwarp();
                        if (++slot == p.nB) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
        if (prof && lane == 0) atomicAdd(&g_tc2_prof[6], (unsigned long long)prof_acc[0]);
    } else if (warp == T2_WARP_RELAY) {
        // ===================== cta_group::2 relay: tells the leader's MMA thread that THIS CTA's operands of a stage have landed ====
        // (the bulk copies signal barriers of their own CTA only; the leader waits for one barrier per stage, count 2)
        if (cg2 && lane == 0) {     // constant-folded away in the cta_group::1 instantiation
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
            for (int tile = tile0; tile < p.n_tiles; tile += tstep) {
                const TileUnits tu = tile_units(p, (int)tile_decode(p, tile).y, (int)crank);
                for (int kf = 0; kf < p.KF; ++kf) {
                    if (!(((tu.any0 | tu.any1) >> kf) & 1u)) continue;
                    for (int g = 0; g < p.G; ++g) {
                        mbar_wait(a_full + sa, pha);
                        mbar_wait(b_full + sb, phb);
                        mbar_arrive_cluster(ready + sb, 0u);
                        if (++sb == p.nB) { sb = 0; phb ^= 1; }
                        if (++sa == p.nA) { sa = 0; pha ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == T2_WARP_MMA) {
        // ===================== MMA issuer: ONE elected thread runs the whole loop =====================
        // Measured with the pipeline profile (AID_TC_DEBUG bit 2048) on the narrow layers: the issuing warp never waited for
        // operands or accumulators (waits < 7 % of its time); it spent 60 % inside the per-stage issue region and 34 % in the
        // code around it (per-tile unit arithmetic, per-stage elect / reconvergence / fences): ~1300 cycles of overhead per
        // 24-MMA stage against ~1100 cycles of MMA issue, with the tensor pipe idle meanwhile (ncu: 30 % active at 64 couts).
        // The whole role is therefore a single-thread region (no per-stage elect / __syncwarp, everything in uniform registers),
        // descriptors are built from 32-bit words inside the asm, and a stage's MMAs are issued by unrolled asm blocks.
        // cta_group::2: only the leader CTA issues (M = 256: unit j of both CTAs per MMA), its commits arrive in both CTAs.
        if ((!cg2 || crank == 0) && elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Ntile >> 3) << 17) | ((cg2 ? 256u >> 4 : 128u >> 4) << 24);  // F16 x F16 -> F32, K-major A/B
            const uint32_t kt_bytes = (uint32_t)p.Ntile * (cg2 ? 64u : 128u);     // one kt sub-tile of this CTA's weight slot
            const uint32_t ringA_u = smem_u32(ringA), ringB_u = smem_u32(ringB);
            const uint32_t acc_stride = (uint32_t)(2 * p.ncol_stride);
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0; int ab = 0; uint32_t aphase = 0;
            const bool prof = (p.dbg & 2048) != 0;
            long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
            // loop invariants of the lean path, in 16-byte descriptor units
            const int nk_last = (p.Cin & 63) ? ((p.Cin & 63) >> 4) : 4;      // 16-channel k-steps of the last group
            const bool lean = p.KT == 3 && p.ktb == 3 && (nk_last == 4 || nk_last == 2) && (p.Cin & 15) == 0 && !(p.dbg & (2 | 32 | 8192));
            const uint32_t adesc = desc_lo_sw128(ringA_u), bdesc = desc_lo_sw128(ringB_u);
            const uint32_t a_slot_d = (uint32_t)a_slot_bytes >> 4, b_slot_d = (uint32_t)p.b_slot_bytes >> 4, ktd = kt_bytes >> 4, unit_d = T2_ASLOT_UNIT >> 4;
            const int KFc = p.KF, KFh = p.KF / 2, Gc = p.G, dilTp = p.dil * p.Tp, nAc = p.nA, nBc = p.nB;
            T2_PROF_T0(t_all);
            for (int tile = tile0; tile < p.n_tiles; tile += tstep) {
                T2_PROF_T0(t_su);
                const TileUnits tu = tile_units(p, (int)tile_decode(p, tile).y, (int)crank);
                const Unit2& u0 = tu.u0; const Unit2& u1 = tu.u1;
                const uint32_t vm0 = tu.any0, vm1 = tu.any1, vany = vm0 | vm1;
                T2_PROF_ADD(5, t_su);
                T2_PROF_T0(t_te);
                if constexpr (cg2) mbar_wait(tmem_empty + ab, aphase ^ 1); else mbar_wait(tmem_empty + ab, aphase ^ 1);
                T2_PROF_ADD(0, t_te);
                tc_fence_after();
                uint32_t acc0 = 0u, acc1 = 0u;      // accumulate flags: 0 for the first MMA of each accumulator
                const uint32_t d0 = tmem_base + (uint32_t)ab * acc_stride, d1 = d0 + (uint32_t)p.ncol_stride;
                if (lean) {
                    // hot configuration (3 kt taps in one weight slot, groups of 4 (or, last group, 2) k-steps): per stage one or two
                    // waits, three descriptor bases, 24 MMAs from unrolled asm blocks, two commits -- every other quantity is loop invariant
                    for (int kf = 0; kf < KFc; ++kf) {
                        if (!((vany >> kf) & 1u)) continue;
                        const bool v0 = (vm0 >> kf) & 1u, v1 = (vm1 >> kf) & 1u;
                        const int ft = (kf - KFh) * dilTp;
                        const uint32_t r0d = (uint32_t)((u0.win_start + ft) & 7) * 8u, r1d = (uint32_t)((u1.win_start + ft) & 7) * 8u + unit_d;
                        for (int g = 0; g < Gc; ++g) {
                            if constexpr (cg2) {
                                T2_PROF_T0(t_af);
                                mbar_wait(ready + sb, phb);
                                T2_PROF_ADD(1, t_af);
                            } else {
                                T2_PROF_T0(t_af);
                                mbar_wait(a_full + sa, pha);
                                T2_PROF_ADD(1, t_af);
                                T2_PROF_T0(t_bf);
                                mbar_wait(b_full + sb, phb);
                                T2_PROF_ADD(2, t_bf);
                            }
                            tc_fence_after();
                            T2_PROF_T0(t_is);
                            const uint32_t ad = adesc + (uint32_t)sa * a_slot_d, b = bdesc + (uint32_t)sb * b_slot_d;
                            const uint32_t a0 = ad + r0d, a1 = ad + r1d;
                            const bool full_g = g + 1 < Gc || nk_last == 4;
                            if constexpr (cg2) {
                                if (full_g) tc_stage3<2, 4>(d0, d1, a0, a1, b, ktd, idesc, v0, v1, acc0, acc1);
                                else tc_stage3<2, 2>(d0, d1, a0, a1, b, ktd, idesc, v0, v1, acc0, acc1);
                                tc_commit_cg2(b_empty + sb);
                                tc_commit_cg2(a_empty + sa);
                            } else {
                                if (full_g) tc_stage3<1, 4>(d0, d1, a0, a1, b, ktd, idesc, v0, v1, acc0, acc1);
                                else tc_stage3<1, 2>(d0, d1, a0, a1, b, ktd, idesc, v0, v1, acc0, acc1);
                                tc_commit(b_empty + sb);
                                tc_commit(a_empty + sa);
                            }
                            if (v0) acc0 = 1u;
                            if (v1) acc1 = 1u;
                            T2_PROF_ADD(3, t_is);
                            if (++sb == nBc) { sb = 0; phb ^= 1; }
                            if (++sa == nAc) { sa = 0; pha ^= 1; }
                        }
                    }
                } else if constexpr (!cg2)
                for (int kf = 0; kf < p.KF; ++kf) {      // general path, cta_group::1 only (the launcher keeps cta_group::2 on the lean one)
                    if (!((vany >> kf) & 1u)) continue;
                    const int foff = (kf - p.KF / 2) * p.dil;
                    const bool v0 = (vm0 >> kf) & 1u, v1 = (vm1 >> kf) & 1u;
                    const uint32_t r0 = (uint32_t)((u0.win_start + foff * p.Tp) & 7), r1 = (uint32_t)((u1.win_start + foff * p.Tp) & 7);
                    for (int g = 0; g < p.G; ++g) {
                        const int nk = min(4, (p.Cin - g * 64) >> 4);     // 16-channel k-steps in this group
                        T2_PROF_T0(t_af);
                        mbar_wait(a_full + sa, pha);
                        T2_PROF_ADD(1, t_af);
                        const uint32_t abase = ringA_u + (uint32_t)sa * (uint32_t)a_slot_bytes;
                        const uint32_t a0row = desc_lo_sw128(abase + r0 * 128u), a1row = desc_lo_sw128(abase + T2_ASLOT_UNIT + r1 * 128u);
                        for (int c = 0; c < nktb; ++c) {
                            const int nkt = min(p.ktb, p.KT - c * p.ktb);
                            T2_PROF_T0(t_bf);
                            mbar_wait(b_full + sb, phb);
                            T2_PROF_ADD(2, t_bf);
                            tc_fence_after();
                            T2_PROF_T0(t_is);
                            const uint32_t bbase = desc_lo_sw128(ringB_u + (uint32_t)sb * (uint32_t)p.b_slot_bytes);
                            if (!(p.dbg & 2)) {
                                for (int k = 0; k < nkt; ++k) {
                                    // one 128-byte pixel row per kt tap (descriptor units of 16 bytes: +8), kt_bytes per weight sub-tile
                                    const uint32_t kt8 = (uint32_t)(c * p.ktb + k + p.kt_shift) * 8u;
                                    const uint32_t blo = bbase + (uint32_t)k * (kt_bytes >> 4);
                                    if (nk == 4) {
                                        if (v0 && v1) { tc_mma_kx2<1, 4>(d0, d1, a0row + kt8, a1row + kt8, blo, idesc, acc0, acc1); acc0 = 1u; acc1 = 1u; }
                                        else if (v0) { tc_mma_k<1, 4>(d0, a0row + kt8, blo, idesc, acc0); acc0 = 1u; }
                                        else if (v1) { tc_mma_k<1, 4>(d1, a1row + kt8, blo, idesc, acc1); acc1 = 1u; }
                                    } else {
                                        for (int j = 0; j < nk; ++j) {
                                            if (v0) { tc_mma_f16_lo(d0, a0row + kt8 + 2u * j, blo + 2u * j, idesc, acc0); acc0 = 1u; }
                                            if (v1) { tc_mma_f16_lo(d1, a1row + kt8 + 2u * j, blo + 2u * j, idesc, acc1); acc1 = 1u; }
                                        }
                                    }
                                }
                            }
                            tc_commit(b_empty + sb);
                            if (c == nktb - 1) tc_commit(a_empty + sa);
                            T2_PROF_ADD(3, t_is);
                            if (++sb == p.nB) { sb = 0; phb ^= 1; }
                        }
                        if (++sa == p.nA) { sa = 0; pha ^= 1; }
                    }
                }
                if constexpr (cg2) tc_commit_cg2(tmem_full + ab); else tc_commit(tmem_full + ab);
                if (++ab == p.acc_bufs) { ab = 0; aphase ^= 1; }
            }
            T2_PROF_ADD(4, t_all);
            if (prof) {
                for (int k = 0; k < 5; ++k) atomicAdd(&g_tc2_prof[k], (unsigned long long)prof_acc[k]);
                atomicAdd(&g_tc2_prof[10], (unsigned long long)prof_acc[5]);
                atomicAdd(&g_tc2_prof[9], 1ull);
            }
        }
        __syncwarp();
    } else if (warp < T2_EPI_WARP0 + T2_EPI_WARPS) {
        // ===================== epilogue: TMEM -> registers -> out = alpha*(acc*gate + R), statistics =====================
        // EW warps: warp e reads TMEM lane quadrant (e & 3) (one pixel per thread) and column group (e >> 2).  The work is a flat
        // sequence of BW-column batches (tile -> unit -> batch); the residual loads of batch k+1 are issued before batch k is
        // processed, so one batch of loads is always in flight, also across the wait for the next tile's accumulator.
        const int e = warp - T2_EPI_WARP0;
        bool fast = false;
        if constexpr (EW == 8) {
            if (p.epi_class) {      // host: launch_conv_tc2
                fast = true;
                EpiArgs ea{p.out, p.R, p.gate, p.gate_bstride, p.alpha, p.stats, p.Ntile, p.n_ntiles};
                Tc2UnitIter it(p, (int)crank, tile0, tstep, (warp & 3) * 32 + lane);
                float* gsm_base = reinterpret_cast<float*>(bar_base + T2_BAR_BYTES);
                double* sacc_base = reinterpret_cast<double*>(bar_base + T2_BAR_BYTES + T2_STAT_GATE_BYTES);
                switch (p.epi_class) {
                    case 1: epilogue_fast<CG2, 16, 2, 8>(ea, it, e, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty); break;
                    case 2: epilogue_fast<CG2, 24, 2, 12>(ea, it, e, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty); break;
                    case 3: epilogue_fast<CG2, 32, 2, 16>(ea, it, e, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty); break;
                    case 4: epilogue_fast<CG2, 32, 2, 32>(ea, it, e, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty); break;
                    default: epilogue_fast<CG2, 32, 4, 32>(ea, it, e, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty); break;
                }
            }
        }
        if (!fast) {
        const int q = warp & 3;
        const int cw = e >> 2;
        const int ncols = p.Ntile / NCW;        // multiple of 8; equals GP statistics groups when Ntile == Ntot
        const int cbeg = cw * ncols;
        const int gcn = p.Ntot / 8;
        const float al = p.alpha, gs = T2_OUT_SCALE * p.alpha;
        const int osc = (int)p.out.sc, rsc = (int)p.R.sc;   // plane strides (< 2^31 elements)
        const bool has_r = p.R.p != nullptr, has_gate = p.gate != nullptr, do_stats = p.stats != nullptr;
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;

        // gate[c] * 2^-14 * alpha of this warp's columns, private copy in shared memory (refreshed when the clip changes)
        float* gsm = reinterpret_cast<float*>(bar_base + T2_BAR_BYTES) + e * GSM_W;
        int gate_key = -2;
        const uint32_t gsm_addr = smem_u32(gsm);
        const int pofs = q * 32 + lane;           // this thread's pixel inside a unit

        // batch descriptors as plain scalars (a struct passed through lambdas ends up in local memory)
        float* c_po = nullptr; uint32_t c_oo = 0; uint32_t c_tcol = 0; int c_ui = 0, c_c0 = 0, c_b = 0, c_ab = 0, c_nt = 0;
        uint32_t c_aphase = 0; bool c_ok = false, c_last = false, c_ulast = false;
        const float* n_pr = nullptr; float* n_po = nullptr; uint32_t n_ro = 0, n_oo = 0; uint32_t n_tcol = 0; int n_ui = 0, n_c0 = 0, n_b = 0, n_ab = 0, n_nt = 0;
        uint32_t n_aphase = 0; bool n_ok = false, n_valid = false, n_last = false, n_ulast = false;
        // iterator state of the next batch to set up; per-unit values are recomputed only at the first batch of a unit
        int it_tile = tile0, it_ui = 0, it_c0 = 0, it_ab = 0; uint32_t it_aphase = 0;
        const float* u_pr = nullptr; float* u_po = nullptr; uint32_t u_ro = 0, u_oo = 0; int u_b = 0, u_nt = 0; bool u_ok = false, u_has1 = false;
        auto setup_next = [&]() {
            n_valid = it_tile < p.n_tiles;
            if (!n_valid) return;
            if (it_c0 == 0) {
                uint32_t pair = (uint32_t)it_tile; u_nt = 0;
                if (p.n_ntiles > 1) { const uint2 dm = tile_decode(p, it_tile); u_nt = (int)dm.x; pair = dm.y; }
                u_has1 = unit_index(p, (int)pair, (int)crank, 1) < p.n_units;
                const int co0 = u_nt * p.Ntile + cbeg;
                Unit2 u = unit2_info(p, unit_index(p, (int)pair, (int)crank, it_ui));
                if (!u.exists) u.b = 0;                   // cta_group::2, last quad: this CTA only keeps the handshakes going
                const int o = u.o0 + pofs;                // output position in the padded stream of the real rows
                const int row = (int)fast_divmod((uint32_t)o, (uint32_t)p.Tp, p.mg_Tp).x, tp = o - row * p.Tp;
                u_ok = u.exists && tp >= 1 && tp <= p.T && row <= u.f_hi;
                // lanes that own no real pixel read (never write) pixel 0 of their clip: the loads need no predicate
                const long long pix = u_ok ? (long long)row * p.T + (tp - 1) : 0;
                // warp-uniform clip base pointers + 32-bit per-thread element offsets (a clip's tensor has < 2^31 elements)
                u_po = p.out.p + (long long)u.b * p.out.sb;
                u_pr = p.R.p + (long long)u.b * p.R.sb;
                u_oo = p.out_cl ? (uint32_t)pix * (uint32_t)p.out.C + (uint32_t)co0 : (uint32_t)co0 * (uint32_t)osc + (uint32_t)pix;
                u_ro = p.r_cl ? (uint32_t)pix * (uint32_t)p.R.C + (uint32_t)co0 : (uint32_t)co0 * (uint32_t)rsc + (uint32_t)pix;
                u_b = u.b;
            }
            n_po = u_po; n_pr = u_pr;
            n_oo = u_oo + (p.out_cl ? (uint32_t)it_c0 : (uint32_t)it_c0 * (uint32_t)osc);
            n_ro = u_ro + (p.r_cl ? (uint32_t)it_c0 : (uint32_t)it_c0 * (uint32_t)rsc);
            n_tcol = tq + (uint32_t)(it_ab * 2 * p.ncol_stride + it_ui * p.ncol_stride + it_c0);
            n_ui = it_ui; n_c0 = it_c0; n_b = u_b; n_nt = u_nt; n_ab = it_ab; n_aphase = it_aphase; n_ok = u_ok;
            n_last = false; n_ulast = false;
            it_c0 += BW;
            if (it_c0 >= ncols) {
                it_c0 = 0; n_ulast = true;
                if (it_ui == 1 || !u_has1) {
                    n_last = true;
                    it_ui = 0; it_tile += tstep;
                    if (++it_ab == p.acc_bufs) { it_ab = 0; it_aphase ^= 1; }
                } else it_ui = 1;
            }
        };
        // (Tried and dropped, measured: pulling the NEXT tile's residual lines into L2 with prefetch.global.L2 at the start of each
        // tile changed nothing at 64 couts and cost 1-3 % on the wider layers -- the epilogue is not bound by the latency of
        // its residual loads.)
        float rr[BW], rn[BW];
        auto load_next = [&]() {
            if (n_valid && has_r && !(p.dbg & 1)) {
                const int nb = ncols - n_c0;      // >= 8, multiple of 8; columns past it are not loaded
                if (p.r_cl) {
#pragma unroll
                    for (int j4 = 0; j4 < BW / 4; ++j4) {
                        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (j4 * 4 < nb) t4 = __ldg(reinterpret_cast<const float4*>(n_pr + n_ro) + j4);
                        rn[j4 * 4 + 0] = t4.x; rn[j4 * 4 + 1] = t4.y; rn[j4 * 4 + 2] = t4.z; rn[j4 * 4 + 3] = t4.w;
                    }
                } else if (nb >= BW) {
#pragma unroll
                    for (int j = 0; j < BW; ++j) rn[j] = n_pr[n_ro + (uint32_t)j * (uint32_t)rsc];   // may alias out: plain loads
                } else {
#pragma unroll
                    for (int j = 0; j < BW; ++j) rn[j] = (j < nb) ? n_pr[n_ro + (uint32_t)j * (uint32_t)rsc] : 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < BW; ++j) rn[j] = 0.f;
            }
        };
        // Statistics: (sum, sumsq) of the 4 groups this warp's columns cover.  A thread sums ONE unit's values of its pixel in
        // fp32 registers (a fixed set of values in a fixed order, whatever the batch size or the tile -> CTA assignment), adds that
        // partial to its private double accumulators in shared memory at the end of the unit, and the doubles are reduced and
        // added to the global accumulators when the clip changes.  Everything order-dependent happens in double on fp32 terms,
        // so the statistics -- and with them the fp16 roundings of the next layer's operands -- are reproducible run to run and
        // identical for a clip evaluated alone or inside a batch (to ~1e-16 relative).
        float SQ[2 * GP];                    // (sum, sumsq) of group k at [2k], [2k + 1]
#pragma unroll
        for (int k = 0; k < 2 * GP; ++k) SQ[k] = 0.f;
        int b_cur = -1, nt_cur = 0, gk = 0, gpos = 0;
        constexpr int SACC_STRIDE = EW * 32;   // [k < 2 GP][epilogue threads]: 16 KB for either EW
        double* sacc = reinterpret_cast<double*>(bar_base + T2_BAR_BYTES + T2_STAT_GATE_BYTES) + (e * 32 + lane);
        if (do_stats) {
#pragma unroll
            for (int k = 0; k < 2 * GP; ++k) sacc[k * SACC_STRIDE] = 0.0;
        }
        auto unit_end = [&]() {
            if (do_stats) {
#pragma unroll
                for (int k = 0; k < 2 * GP; ++k) sacc[k * SACC_STRIDE] += (double)SQ[k];
            }
#pragma unroll
            for (int k = 0; k < 2 * GP; ++k) SQ[k] = 0.f;
        };
        auto flush_stats = [&]() {
            if (do_stats && b_cur >= 0) {
                double v[2 * GP];
#pragma unroll
                for (int k = 0; k < 2 * GP; ++k) {
                    v[k] = sacc[k * SACC_STRIDE]; sacc[k * SACC_STRIDE] = 0.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
                }
                if (lane == 0) {
                    const int g0 = (nt_cur * p.Ntile + cbeg) / gcn;
#pragma unroll
                    for (int k = 0; k < 2 * GP; ++k)
                        if (g0 + (k >> 1) < 8) atomicAdd(p.stats + (long long)b_cur * 16 + (g0 + (k >> 1)) * 2 + (k & 1), v[k]);
                }
            }
        };
        auto add_group = [&](float s, float qq) {
#pragma unroll
            for (int k = 0; k < GP; ++k)
                if (gk == k) { SQ[2 * k] += s; SQ[2 * k + 1] += qq; }
        };
        // one 8-column chunk: out = acc*gate' + R*alpha, store, statistics
        auto chunk = [&](const uint32_t* acc8, const float* r8, uint32_t g8, uint32_t oo8) {
            float g[8];
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g[0]), "=f"(g[1]), "=f"(g[2]), "=f"(g[3]) : "r"(g8));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(g[4]), "=f"(g[5]), "=f"(g[6]), "=f"(g[7]) : "r"(g8));
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(acc8[j]), g[j], r8[j] * al);
            if (c_ok && !(p.dbg & 256)) {
                if (p.out_cl) {
                    float4* o4 = reinterpret_cast<float4*>(c_po + oo8);
                    o4[0] = make_float4(v[0], v[1], v[2], v[3]);
                    o4[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) c_po[oo8 + (uint32_t)j * (uint32_t)osc] = v[j];
                }
            }
            if (do_stats) {
                const float m = c_ok ? 1.f : 0.f;
                if (gpos + 8 <= gcn) {
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s += v[j]; qq = fmaf(v[j], v[j], qq); }
                    add_group(s * m, qq * m);
                    gpos += 8;
                    if (gpos == gcn) { gpos = 0; ++gk; }
                } else if (gcn >= 8) {
                    // one group boundary inside the chunk (group widths that are not multiples of 8, e.g. 12 for 96 channels)
                    const int first = gcn - gpos;       // columns [0, first) finish the current group
                    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < first) { s0 += v[j]; q0 = fmaf(v[j], v[j], q0); } else { s1 += v[j]; q1 = fmaf(v[j], v[j], q1); }
                    }
                    add_group(s0 * m, q0 * m); ++gk;
                    add_group(s1 * m, q1 * m); gpos = 8 - first;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {     // narrow test networks only (groups of 2, 4 or 6 channels)
                        add_group(v[j] * m, v[j] * v[j] * m);
                        if (++gpos == gcn) { gpos = 0; ++gk; }
                    }
                }
            }
        };

        const bool prof = (p.dbg & 2048) != 0 && e == 0;
        long long prof_acc[2] = {0, 0};
        T2_PROF_T0(t_all);
        setup_next();
        load_next();
        while (n_valid) {
            // the prefetched batch becomes the current one
            c_po = n_po; c_oo = n_oo; c_tcol = n_tcol; c_ui = n_ui; c_c0 = n_c0; c_b = n_b; c_ab = n_ab; c_nt = n_nt; c_aphase = n_aphase;
            c_ok = n_ok; c_last = n_last; c_ulast = n_ulast;
#pragma unroll
            for (int j = 0; j < BW; ++j) rr[j] = rn[j];
            setup_next();
            load_next();
            if (c_ui == 0 && c_c0 == 0) { T2_PROF_T0(t_w); mbar_wait(tmem_full + c_ab, c_aphase); T2_PROF_ADD(0, t_w); tc_fence_after(); }
            if (c_c0 == 0) {
                gk = 0; gpos = 0;
                if (c_b != b_cur || c_nt != nt_cur) { flush_stats(); b_cur = c_b; nt_cur = c_nt; }
                const int gkey = p.gate_bstride ? c_b * p.n_ntiles + c_nt : c_nt;
                if (gkey != gate_key) {
                    gate_key = gkey;
                    __syncwarp();
                    for (int k = lane; k < ncols; k += 32)
                        gsm[k] = has_gate ? __ldg(p.gate + (long long)c_b * p.gate_bstride + c_nt * p.Ntile + cbeg + k) * gs : gs;
                    __syncwarp();
                }
            }
            if (!(p.dbg & 1)) {
                uint32_t acc[BW];
                if (!(p.dbg & 512)) {
                    if constexpr (BW == 32) tmem_ld32_nowait(c_tcol, acc); else tmem_ld16_nowait(c_tcol, acc);   // columns past this warp's range are read but never used
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < BW; ++j) acc[j] = 0x3f800000u;
                }
                const int nb = ncols - c_c0;
                const uint32_t gq = gsm_addr + (uint32_t)c_c0 * 4u;
                if (nb >= BW) {
#pragma unroll
                    for (int j8 = 0; j8 < BW / 8; ++j8) chunk(acc + j8 * 8, rr + j8 * 8, gq + j8 * 32, c_oo + (p.out_cl ? (uint32_t)(j8 * 8) : (uint32_t)(j8 * 8) * (uint32_t)osc));
                } else {
#pragma unroll
                    for (int j8 = 0; j8 < BW / 8 - 1; ++j8)
                        if (j8 * 8 < nb) chunk(acc + j8 * 8, rr + j8 * 8, gq + j8 * 32, c_oo + (p.out_cl ? (uint32_t)(j8 * 8) : (uint32_t)(j8 * 8) * (uint32_t)osc));
                }
            }
            if (c_ulast) unit_end();
            if (c_last) {      // one arrival per warp: 256 per-thread arrivals on one mbarrier serialise in the shared-memory pipe
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (cg2) mbar_arrive_cluster(tmem_empty + c_ab, 0u); else mbar_arrive(tmem_empty + c_ab); }   // the leader's MMA thread waits
            }
        }
        flush_stats();
        T2_PROF_ADD(1, t_all);
        if (prof && lane == 0) { atomicAdd(&g_tc2_prof[7], (unsigned long long)prof_acc[0]); atomicAdd(&g_tc2_prof[8], (unsigned long long)prof_acc[1]); }
        }   // generic epilogue
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (cg2) cluster_sync_all();   // the partner may still signal this CTA's barriers / read its shared memory until it is done too
    if (warp == T2_WARP_MMA) {
        tc_fence_after();
        if constexpr (cg2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int EW> __global__ void __launch_bounds__((EW + 4) * 32, 1) conv_tc2_kernel(const __grid_constant__ Tc2Args p) { conv_tc2_body<false, EW>(p); }
template <int EW> __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((EW + 4) * 32, 1) conv_tc2_cg2_kernel(const __grid_constant__ Tc2Args p) { conv_tc2_body<true, EW>(p); }

// ---- operand preparation ---------------------------------------------------------------------------------
// Couts per tile.  256-wide tiles fill TMEM with one accumulator (2 units x 256 columns), so their epilogue cannot overlap the
// next tile's MMAs; multi-tap convolutions with 256 couts run as two 128-wide n-tiles with double-buffered accumulators
// instead (AID_TC2_SPLIT256=0 restores the single tile).  1x1 convolutions are HBM bound and keep the widest tile.
static int tc2_ntile(int Cout, int taps) {
    static const int split256 = getenv("AID_TC2_SPLIT256") ? atoi(getenv("AID_TC2_SPLIT256")) : 1;
    if (Cout == 256 && taps > 1 && split256) return 128;
    return Cout <= 256 ? Cout : 256;
}

size_t tc2_weight_halves(int Cout, int Cin, int KF, int KT) { return (size_t)Cout * KF * KT * ((Cin + 63) / 64) * 64; }

// w[co][ci][kf][kt] (fp32) -> [n-tile][kf][G][kt][Ntile][64] fp16 (x 2^10), 16-byte chunks swizzled by (n & 7); channels past Cin = 0
__global__ void pack_weight_tc2_kernel(const float* __restrict__ w, __half* __restrict__ wp, int Ntot, int Ntile, int Cin, int KF, int KT,
                                       unsigned long long* __restrict__ sat) {
    const int G = (Cin + 63) / 64;
    const long long total = (long long)Ntot * KF * KT * G * 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int pos = (int)(r % 64); r /= 64;       // physical position inside the 128-byte row
        const int n = (int)(r % Ntile); r /= Ntile;
        const int kt = (int)(r % KT); r /= KT;
        const int g = (int)(r % G); r /= G;
        const int kf = (int)(r % KF); r /= KF;
        const int nt = (int)r;
        const int chunk = (pos >> 3) ^ (n & 7);       // logical chunk stored at this position
        const int ci = g * 64 + chunk * 8 + (pos & 7), co = nt * Ntile + n;
        float v = 0.f;
        if (ci < Cin) v = w[(((long long)co * Cin + ci) * KF + kf) * KT + kt] * T2_W_SCALE;
        if (sat && fabsf(v) > 60000.f) atomicAdd(sat, 1ull);
        v = fminf(fmaxf(v, -60000.f), 60000.f);
        wp[i] = __float2half_rn(v);
    }
}

void launch_pack_weight_tc2(const float* w, __half* wp, int Cout, int Cin, int KF, int KT, cudaStream_t s, unsigned long long* sat) {
    const long long total = (long long)tc2_weight_halves(Cout, Cin, KF, KT);
    pack_weight_tc2_kernel<<<(int)min((long long)8192, (total + 255) / 256), 256, 0, s>>>(w, wp, Cout, tc2_ntile(Cout, KF * KT), Cin, KF, KT, sat);
    AID_COUNT_LAUNCH(1);
}

// 16 * GELU(v) with the exact-erf definition to ~5e-7 absolute (Abramowitz & Stegun 7.1.26, MUFU rcp / ex2): erff() costs
// about twice the instructions, and this HBM-streaming pass was instruction bound; the result is rounded to fp16 (2^-11
// relative) right after, so the approximation error is invisible.  The operand scale (x16) is folded into the last FMA.
__device__ __forceinline__ float gelu16_tc2(float v) {
    const float ax = fabsf(v) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
    float pl = fmaf(t, 1.061405429f, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    pl *= t;
    const float z = v * 0.84932180028801904272f;          // sqrt(log2(e) / 2): exp(-v^2 / 2) = 2^(-z^2)
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-z * z));
    const float erf_abs = fmaf(-pl, ex, 1.f);
    const float h = 8.f * v;
    return fmaf(fabsf(h), erf_abs, h);                    // h * (1 + sign(v) * erf(|v| / sqrt 2))
}
// The same value for v = x * s with the per-channel scale folded into two constants, cu = |s| * sqrt(log2(e) / 2) and
// ch = 8 s: u = |x| cu serves both the exponent (2^(-u^2)) and the rational argument (|v| / sqrt 2 = u / sqrt(log2 e)),
// 14 instructions per element instead of 16.
__device__ __forceinline__ float gelu16_tc2_folded(float x, float cu, float ch) {
    const float u = fabsf(x) * cu;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.83255461115769775635f, u, 1.f)));
    float pl = fmaf(t, 1.061405429f, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    pl *= t;
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-u * u));
    const float erf_abs = fmaf(-pl, ex, 1.f);
    const float h = x * ch;
    return fmaf(fabsf(h), erf_abs, h);
}
__device__ __forceinline__ bool ld_ok(int i, int n) { return i < n; }
// Normalise / modulate / GELU (unet.py:159-163, 479, 482) writing the channels-last fp16 operand:
//   a[b][g][r][tp][chunk ^ ((r*Tp + tp) & 7)][8] = fp16(16 * act(x[b, 64g + 8 chunk + j, r - PF, tp - 1] * scale_c)),
//   pad pixels (tp = 0, T+1), pad rows and channels past C = 0.  stats == nullptr: plain layout/precision conversion.
// The pass is instruction-issue bound, not HBM bound (17 instructions of GELU per element), so the layout work around it is
// kept to a few instructions per element: the source plane [F][T] of a channel is a flat run (rows are contiguous), a block
// converts 128 consecutive source pixels per iteration whatever the row length, thread (warp w, lane) loads 4 consecutive
// pixels of the channels [8w, 8w+8) of the group with 8 x LDG.128 and leaves four 16-byte chunks in a shared tile whose
// private layout (chunk slot = (chunk + pixel / 4) & 7) is conflict free for both sides; the way out applies the operand
// swizzle, writes 128 contiguous bytes per pixel and the zero pad pixels next to the first / last pixel of a row.
// grid: (B * G, chunk shares), block 256.  vec == 0 (T or the view not 16-byte aligned): scalar loads, same mapping.
// UP (plain conversion only): the operand's first up.C channels are not read from x but computed on the fly as the 2x time-upsampling
// (unet.py:128-150, the arithmetic of resample_up4_kernel) of up -- a [B][up.C][F][T / 2] tensor -- and x supplies the channels behind them:
// the decoder's concatenation [upsampled decoder stream | encoder skip] is consumed only through this operand (proj_in / res_conv of the
// level's main block), so its upsampled half never exists in fp32.
struct UpSrc { const float* p; long long sb, sc; int C, T; };
template <bool COUNT, bool UP>
__global__ void __launch_bounds__(256, 3)
gn_act_tc2_kernel(TV x, const double* __restrict__ stats, double n_per_group, const float* __restrict__ gamma,
                  const float* __restrict__ affine, long long affine_bstride, int gelu, int PF, int G, int vec, uint32_t mg_T,
                  __half* __restrict__ a, unsigned long long* __restrict__ sat, UpSrc up) {
    unsigned int nsat = 0;   // COUNT: values this thread clamped to the finite fp16 range (aid_debug_saturation)
    const int T = x.T, Tp = T + 2, rows_total = x.F + 2 * PF;
    const int n_src = x.F * T;                        // source pixels of one channel plane
    const int g = blockIdx.x % G, b = blockIdx.x / G;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int upC = UP ? up.C : 0, Ctot = x.C + upC;  // operand channels: [0, upC) upsampled from up, [upC, Ctot) from x
    __shared__ float s_scale[64];
    __shared__ float s_inv[8];
    __shared__ __align__(16) uint8_t tile[128 * 128];   // [pixel][slot][16 B]
    // 1 / (unbiased std + eps) of the 8 statistics groups in double (8 threads), then the 64 per-channel scales of this group
    if (threadIdx.x < 8 && stats) {
        const double s1 = stats[((long long)b * 8 + threadIdx.x) * 2 + 0], s2 = stats[((long long)b * 8 + threadIdx.x) * 2 + 1];
        double var = (s2 - s1 * s1 / n_per_group) / (n_per_group - 1.0);
        var = var > 0.0 ? var : 0.0;
        s_inv[threadIdx.x] = 1.f / ((float)sqrt(var) + 1e-7f);
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        float sc = 0.f;
        const int c = g * 64 + threadIdx.x;
        if (c < Ctot) {
            sc = 1.f;
            if (stats) {
                const float mod = affine ? (1.f + affine[b * affine_bstride + c]) : 1.f;
                sc = gamma[c] * mod * s_inv[c / (x.C / 8)];
            } else {
                // plain conversion (no normalisation): optional per-channel vector gamma[b * affine_bstride + c] and per-tensor
                // device scalar affine[0] (the data-gradient path scales its operand into the fp16 range this way)
                if (gamma) sc = gamma[b * affine_bstride + c];
                if (affine) sc *= affine[0];
            }
        }
        s_scale[threadIdx.x] = sc;
    }
    __syncthreads();
    float cu[8], chh[8];                         // gelu: |scale| sqrt(log2(e) / 2) and 8 scale; otherwise chh = 16 scale
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float sj = s_scale[w * 8 + j];
        cu[j] = fabsf(sj) * 0.84932180028801904272f;
        chh[j] = gelu ? 8.f * sj : T2_A_SCALE * sj;
    }
    const int c0 = g * 64 + w * 8;               // first channel of this warp's chunk
    const bool chok = c0 < Ctot;                 // C is a multiple of 8: a chunk is either all real or all padding
    const bool from_up = UP && c0 < upC;         // (warp-uniform; upC is a multiple of 8)
    uint8_t* dst = reinterpret_cast<uint8_t*>(a + ((long long)b * G + g) * rows_total * Tp * 64);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    if (PF > 0) {                                // zero rows above and below the plane, shared by the blocks of the plane
        uint4* top = reinterpret_cast<uint4*>(dst);
        uint4* bot = reinterpret_cast<uint4*>(dst + (long long)(x.F + PF) * Tp * 128);
        const int n16 = PF * Tp * 8;
        for (int i = blockIdx.y * 256 + threadIdx.x; i < n16; i += gridDim.y * 256) { top[i] = zero4; bot[i] = zero4; }
    }
    const float* src = x.p + (long long)b * x.sb + (long long)(c0 - upC) * x.sc + 4 * lane;
    const int nchunk = (n_src + 127) >> 7;
    const int slot_w = ((w + lane) & 7) << 4;    // this thread's pixels are 4 lane + i: pixel / 4 == lane
    const int cp = threadIdx.x & 7;              // way out: 16-byte position inside the 128-byte pixel row
    uint8_t* dst_cp = dst + cp * 16;
    for (int ch = blockIdx.y; ch < nchunk; ch += gridDim.y) {
        const int s0 = ch << 7, st = s0 + 4 * lane;
        float v[8][4];
        if (UP && from_up) {
            // 4 outputs m = 4v .. 4v+3 of row f from the 6 inputs x[2v-2 .. 2v+3] of the half-rate row (reflect padded); T % 4 == 0.
            // A lane loads its own pair x[2v], x[2v+1] (one coalesced LDG.64 per channel) and takes the pairs left and right of it from
            // its neighbour lanes; the lanes at a row end or at the edge of the warp's 128-pixel run load theirs (six scalar loads
            // per channel and lane cost six 256-byte L1 accesses per warp instead of one).
            const bool act = st < n_src;
            const uint2 ft = fast_divmod((uint32_t)(act ? st : 0), (uint32_t)T, mg_T);
            const int vq = (int)ft.y >> 2, Th = up.T;
            const float* row = up.p + (long long)b * up.sb + (long long)c0 * up.sc + (long long)ft.x * Th;
            const bool own_l = act && (lane == 0 || vq == 0), own_r = act && (lane == 31 || vq == (T >> 2) - 1);
            int il0 = 2 * vq - 2, il1 = 2 * vq - 1, ir0 = 2 * vq + 2, ir1 = 2 * vq + 3;       // reflect: -n for n < 0, 2 (Th - 1) - n for n >= Th
            il0 = il0 < 0 ? -il0 : il0; il1 = il1 < 0 ? -il1 : il1;
            ir0 = ir0 >= Th ? 2 * (Th - 1) - ir0 : ir0; ir1 = ir1 >= Th ? 2 * (Th - 1) - ir1 : ir1;
            constexpr float kc[8] = {-0.01171875f, -0.03515625f, 0.11328125f, 0.43359375f, 0.43359375f, 0.11328125f, -0.03515625f, -0.01171875f};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* rj = row + (long long)j * up.sc;
                float2 mine = make_float2(0.f, 0.f);
                if (act) mine = __ldg(reinterpret_cast<const float2*>(rj + 2 * vq));
                float wv[6];
                wv[0] = __shfl_up_sync(0xffffffffu, mine.x, 1); wv[1] = __shfl_up_sync(0xffffffffu, mine.y, 1);
                wv[4] = __shfl_down_sync(0xffffffffu, mine.x, 1); wv[5] = __shfl_down_sync(0xffffffffu, mine.y, 1);
                wv[2] = mine.x; wv[3] = mine.y;
                if (own_l) { wv[0] = __ldg(rj + il0); wv[1] = __ldg(rj + il1); }
                if (own_r) { wv[4] = __ldg(rj + ir0); wv[5] = __ldg(rj + ir1); }
                float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    y[0] += kc[7 - 2 * q] * wv[q];
                    y[1] += kc[6 - 2 * q] * wv[q + 1];
                    y[2] += kc[7 - 2 * q] * wv[q + 1];
                    y[3] += kc[6 - 2 * q] * wv[q + 2];
                }
                v[j][0] = act ? y[0] : 0.f; v[j][1] = act ? y[1] : 0.f; v[j][2] = act ? y[2] : 0.f; v[j][3] = act ? y[3] : 0.f;
            }
        } else if (chok) {
            if (vec) {
                const bool ld = st < n_src;      // n_src is a multiple of 4 here
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 q = ld ? __ldg(reinterpret_cast<const float4*>(src + (long long)j * x.sc + s0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[j][0] = q.x; v[j][1] = q.y; v[j][2] = q.z; v[j][3] = q.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[j][i] = st + i < n_src ? __ldg(src + (long long)j * x.sc + s0 + i) : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint4 hv = zero4;
            if (chok) {
                float r[8];
                if (gelu) {
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const float2 r2 = gelu16_tc2_folded2(make_float2(v[j][i], v[j + 1][i]), make_float2(cu[j], cu[j + 1]), make_float2(chh[j], chh[j + 1]));
                        r[j] = r2.x; r[j + 1] = r2.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = v[j][i] * chh[j];
                }
                if (COUNT) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) nsat += (!(fabsf(r[j]) <= 65504.f) && (ld_ok(st + i, n_src))) ? 1u : 0u;
                }
                hv = make_uint4(pack_half2_sat(r[0], r[1]), pack_half2_sat(r[2], r[3]), pack_half2_sat(r[4], r[5]), pack_half2_sat(r[6], r[7]));
            }
            *reinterpret_cast<uint4*>(tile + (4 * lane + i) * 128 + slot_w) = hv;
        }
        __syncthreads();
        const int p0 = threadIdx.x >> 3;
        if (T % 32 == 0) {
            // the 32 pixels of a pass share a row; row / column of the first pass once, then incrementally (block-uniform)
            const uint2 ft = fast_divmod((uint32_t)s0, (uint32_t)T, mg_T);
            int tk = (int)ft.y, qk = ((int)ft.x + PF) * Tp + tk + 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (s0 + 32 * k < n_src) {                              // n_src is a multiple of 32: a pass is all in or all out
                    const int p = p0 + 32 * k, q = qk + p0;
                    const int c = cp ^ (q & 7);                         // the chunk that lives at position cp of that pixel row
                    uint8_t* o = dst_cp + (long long)q * 128;
                    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(tile + p * 128 + (((c + (p >> 2)) & 7) << 4));
                    if (tk == 0 && p0 == 0) *reinterpret_cast<uint4*>(o - 128) = zero4;
                    if (tk + 32 == T && p0 == 31) *reinterpret_cast<uint4*>(o + 128) = zero4;
                }
                tk += 32; qk += 32;
                if (tk == T) { tk = 0; qk += 2; }                       // next row: two pad pixels in between
            }
        } else {
            int p = p0;
            const uint2 ft = fast_divmod((uint32_t)(s0 + p), (uint32_t)T, mg_T);
            int t = (int)ft.y, q = ((int)ft.x + PF) * Tp + t + 1;      // column, flattened padded pixel index
            for (int k = 0; k < 4; ++k) {
                if (s0 + p < n_src) {
                    const int c = cp ^ (q & 7);
                    uint8_t* o = dst_cp + (long long)q * 128;
                    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(tile + p * 128 + (((c + (p >> 2)) & 7) << 4));
                    if (t == 0) *reinterpret_cast<uint4*>(o - 128) = zero4;
                    if (t == T - 1) *reinterpret_cast<uint4*>(o + 128) = zero4;
                }
                p += 32; t += 32; q += 32;
                while (t >= T) { t -= T; q += 2; }                      // next row: two pad pixels in between
            }
        }
        __syncthreads();
    }
    if (COUNT && nsat) atomicAdd(sat, (unsigned long long)nsat);
}

// Same operand, from a channels-last input x[B][F][T][C] (the residual stream inside a conv_mode 2 residual block).  Pure
// streaming: 8 consecutive lanes read one pixel's 64-channel group (256 contiguous bytes), each thread converts one 8-channel
// chunk and writes its 16 bytes at the swizzled position of the same pixel row.  grid: (B * G * rows_total, segment chunks)
__global__ void __launch_bounds__(256)
gn_act_tc2_cl_kernel(const float* __restrict__ x, int B, int C, int F, int T, const double* __restrict__ stats, double n_per_group,
                     const float* __restrict__ gamma, const float* __restrict__ affine, long long affine_bstride, int gelu, int PF, int G,
                     __half* __restrict__ a) {
    const int Tp = T + 2, rows_total = F + 2 * PF;
    int bid = blockIdx.x;
    const int fr = bid % rows_total; bid /= rows_total;
    const int g = bid % G, b = bid / G;
    __shared__ float s_scale[64];
    if (threadIdx.x < 64) {
        float sc = 0.f;
        const int c = g * 64 + threadIdx.x;
        if (c < C) {
            sc = 1.f;
            if (stats) {
                const int grp = c / (C / 8);
                const double s1 = stats[((long long)b * 8 + grp) * 2 + 0], s2 = stats[((long long)b * 8 + grp) * 2 + 1];
                double var = (s2 - s1 * s1 / n_per_group) / (n_per_group - 1.0);
                var = var > 0.0 ? var : 0.0;
                const float stdv = (float)sqrt(var);
                const float mod = affine ? (1.f + affine[b * affine_bstride + c]) : 1.f;
                sc = gamma[c] * mod / (stdv + 1e-7f);
            }
        }
        s_scale[threadIdx.x] = sc;
    }
    __syncthreads();
    const int chunk = threadIdx.x & 7, psub = threadIdx.x >> 3;   // 32 pixels per pass
    float sc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = s_scale[chunk * 8 + j];
    const int f = fr - PF;
    const bool rowok = f >= 0 && f < F;
    const bool chok = g * 64 + chunk * 8 < C;
    const float* src = x + (((long long)b * F + (rowok ? f : 0)) * T) * C + g * 64 + chunk * 8;
    uint8_t* dst_row = reinterpret_cast<uint8_t*>(a + ((((long long)b * G + g) * rows_total + fr) * Tp) * 64);
    const long long gp_row = (long long)fr * Tp;
    for (int tp = blockIdx.y * 32 + psub; tp < Tp; tp += gridDim.y * 32) {
        const int t = tp - 1;
        uint4 hv = make_uint4(0u, 0u, 0u, 0u);
        if (rowok && chok && t >= 0 && t < T) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + (long long)t * C));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + (long long)t * C) + 1);
            float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] *= sc[j];
                v[j] = gelu ? gelu16_tc2(v[j]) : v[j] * T2_A_SCALE;
            }
            hv = make_uint4(pack_half2_sat(v[0], v[1]), pack_half2_sat(v[2], v[3]), pack_half2_sat(v[4], v[5]), pack_half2_sat(v[6], v[7]));
        }
        const int phase = (int)((gp_row + tp) & 7);
        *reinterpret_cast<uint4*>(dst_row + (long long)tp * 128 + ((chunk ^ phase) << 4)) = hv;
    }
}

// tuning: read and clear the pipeline profile counters (AID_TC_DEBUG bit 2048)
void tc2_read_profile(unsigned long long* out16) {
    AID_CUDA_CHECK(cudaDeviceSynchronize());
    AID_CUDA_CHECK(cudaMemcpyFromSymbol(out16, g_tc2_prof, 16 * sizeof(unsigned long long)));
    unsigned long long z[16] = {0};
    AID_CUDA_CHECK(cudaMemcpyToSymbol(g_tc2_prof, z, sizeof z));
}

size_t tc2_act_halves(int B, int C, int F, int T, int PF) { return (size_t)B * ((C + 63) / 64) * 64 * (F + 2 * PF) * (T + 2); }

void launch_gn_act_tc2(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                       long long affine_bstride, bool gelu, int PF, __half* a, cudaStream_t s, unsigned long long* sat) {
    const int rows_total = x.F + 2 * PF, Tp = x.T + 2, G = (x.C + 63) / 64;
    static const int env_bps = getenv("AID_GN_BPS") ? atoi(getenv("AID_GN_BPS")) : 16;
    if ((long long)rows_total * Tp >= (1ll << 31) / 128)
        throw CudaError(cudaErrorInvalidValue, "gn_act_tc2: plane too large for 32-bit pixel indices", __FILE__, __LINE__);
    const int nchunk = (x.F * x.T + 127) / 128;
    // 3 blocks per SM are resident; env_bps blocks per SM stride over the chunks of their plane (finer shares balance better)
    const long long planes = (long long)x.B * G;
    const int shares = (int)std::min<long long>(nchunk, std::max<long long>(1, ((long long)device_sm_count() * env_bps + planes - 1) / planes));
    const int vec = (x.T % 4 == 0 && x.sb % 4 == 0 && x.sc % 4 == 0 && (reinterpret_cast<uintptr_t>(x.p) & 15) == 0) ? 1 : 0;
    dim3 grid((unsigned)planes, shares);
    if (sat) gn_act_tc2_kernel<true, false><<<grid, 256, 0, s>>>(x, stats, (double)n_per_group, gamma, affine, affine_bstride, gelu ? 1 : 0, PF, G, vec,
                                                              div_magic((uint32_t)x.T), a, sat, UpSrc{});
    else gn_act_tc2_kernel<false, false><<<grid, 256, 0, s>>>(x, stats, (double)n_per_group, gamma, affine, affine_bstride, gelu ? 1 : 0, PF, G, vec,
                                                              div_magic((uint32_t)x.T), a, nullptr, UpSrc{});
    AID_COUNT_LAUNCH(1);
}

// Operand of the concatenation [upsample2x(up) | x] (channels up.C + x.C) without materialising the upsampled half: see UpSrc.
// up: [B][Cu][F][T / 2], x: [B][Cx][F][T]; Cu % 8 == 0, T % 4 == 0, x 16-byte aligned.
bool to_planar_tc2_up_supported(const TV& up, const TV& x) {
    return up.B == x.B && up.F == x.F && 2 * up.T == x.T && up.C % 8 == 0 && x.C % 8 == 0 && x.T % 4 == 0 && up.T >= 4 && x.sb % 4 == 0 && x.sc % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(x.p) & 15) == 0 && (reinterpret_cast<uintptr_t>(up.p) & 7) == 0 && up.sb % 2 == 0 && up.sc % 2 == 0;
}
void launch_to_planar_tc2_up(const TV& up, const TV& x, int PF, __half* a, cudaStream_t s) {
    if (!to_planar_tc2_up_supported(up, x)) throw CudaError(cudaErrorInvalidValue, "to_planar_tc2_up: unsupported shape", __FILE__, __LINE__);
    const int rows_total = x.F + 2 * PF, Tp = x.T + 2, G = (x.C + up.C + 63) / 64;
    static const int env_bps = getenv("AID_GN_BPS") ? atoi(getenv("AID_GN_BPS")) : 16;
    if ((long long)rows_total * Tp >= (1ll << 31) / 128)
        throw CudaError(cudaErrorInvalidValue, "to_planar_tc2_up: plane too large for 32-bit pixel indices", __FILE__, __LINE__);
    const int nchunk = (x.F * x.T + 127) / 128;
    const long long planes = (long long)x.B * G;
    const int shares = (int)std::min<long long>(nchunk, std::max<long long>(1, ((long long)device_sm_count() * env_bps + planes - 1) / planes));
    gn_act_tc2_kernel<false, true><<<dim3((unsigned)planes, shares), 256, 0, s>>>(x, nullptr, 1.0, nullptr, nullptr, 0, 0, PF, G, 1, div_magic((uint32_t)x.T), a, nullptr,
                                                                                UpSrc{up.p, up.sb, up.sc, up.C, up.T});
    AID_COUNT_LAUNCH(1);
}

// x_cl: channels-last fp32 [B][F][T][C]
void launch_gn_act_tc2_cl(const float* x_cl, int B, int C, int F, int T, const double* stats, long long n_per_group, const float* gamma,
                          const float* affine, long long affine_bstride, bool gelu, int PF, __half* a, cudaStream_t s) {
    const int rows_total = F + 2 * PF, Tp = T + 2, G = (C + 63) / 64;
    const int npass = (Tp + 31) / 32;
    const long long rows = (long long)B * G * rows_total;
    int ychunks = 1;
    while (ychunks < npass && rows * ychunks < device_sm_count() * 16) ychunks <<= 1;
    ychunks = min(ychunks, npass);
    dim3 grid((unsigned)rows, ychunks);
    gn_act_tc2_cl_kernel<<<grid, 256, 0, s>>>(x_cl, B, C, F, T, stats, (double)n_per_group, gamma, affine, affine_bstride, gelu ? 1 : 0, PF, G, a);
    AID_COUNT_LAUNCH(1);
}

void launch_to_planar_tc2(const TV& x, int PF, __half* a, cudaStream_t s, unsigned long long* sat) {
    launch_gn_act_tc2(x, nullptr, 1, nullptr, nullptr, 0, false, PF, a, s, sat);
}

// a: [B][ceil(Cin/64)][F + 2*PF][T+2][64] with PF >= tc_pad_rows(T, KF, dil); wp from launch_pack_weight_tc2
void launch_conv_tc2(const __half* a, int PF, const __half* wp, int B, int Cin, int F, int T, int KF, int KT, int dil,
                     const TV& out, const ConvEpilogue& ep, int num_sms, cudaStream_t s) {
    if (PF < tc_pad_rows(T, KF, dil)) throw CudaError(cudaErrorInvalidValue, "conv_tc2: not enough pad rows", __FILE__, __LINE__);
    if (ep.R2.p) throw CudaError(cudaErrorInvalidValue, "conv_tc2: R2 is not supported", __FILE__, __LINE__);
    if (!conv_tc_supported(Cin, out.C, KF, KT)) throw CudaError(cudaErrorInvalidValue, "conv_tc2: unsupported shape", __FILE__, __LINE__);
    Tc2Args p{};
    p.a = a; p.w = wp; p.out = out; p.R = ep.R; p.gate = ep.gate; p.gate_bstride = ep.gate_bstride;
    p.alpha = ep.alpha; p.stats = ep.stats; p.out_cl = ep.out_cl ? 1 : 0; p.r_cl = ep.R_cl ? 1 : 0;
    p.B = B; p.Cin = Cin; p.G = (Cin + 63) / 64; p.Ntot = out.C; p.Ntile = tc2_ntile(out.C, KF * KT); p.n_ntiles = out.C / p.Ntile;
    p.F = F; p.T = T; p.Tp = T + 2; p.dil = dil;
    p.KF = KF; p.KT = KT; p.kt_shift = (KT == 1) ? 1 : 0;
    p.PF = PF; p.rows_total = F + 2 * PF;
    p.stream = (T % 128 != 0) ? 1 : 0;
    p.tiles_t = (T + 127) / 128;
    p.units_per_b = p.stream ? (F * p.Tp + 127) / 128 : F * p.tiles_t;
    p.n_units = B * p.units_per_b;
    p.n_pairs = (p.n_units + 1) / 2;
    p.n_tiles = p.n_pairs * p.n_ntiles;
    static const int env_ntm = getenv("AID_TC2_NT_MINOR") ? atoi(getenv("AID_TC2_NT_MINOR")) : 1;
    p.nt_minor = (p.n_ntiles == 2 && env_ntm) ? 1 : 0;
    p.mg_pairs = div_magic(p.n_pairs); p.mg_upb = div_magic(p.units_per_b); p.mg_Tp = div_magic(p.Tp); p.mg_tt = div_magic(p.tiles_t); p.mg_F = div_magic(p.F);
    static const int env_ktb = getenv("AID_TC2_KTB") ? atoi(getenv("AID_TC2_KTB")) : 0;
    static const int env_nA = getenv("AID_TC2_NA") ? atoi(getenv("AID_TC2_NA")) : 3;
    const int kt_bytes = p.Ntile * 128;
    p.ktb = env_ktb > 0 ? min(env_ktb, KT) : (kt_bytes * KT <= 48 * 1024 ? KT : 1);
    // cta_group::2 (AID_TC2_CG2=0 turns it off): an SS-mode MMA re-reads its A tile (128 x 16) and B tile (N x 16) from shared
    // memory, 6 KB per 32 tensor cycles at N = 64 and 8 KB per 64 at N = 128 next to the bulk-copy writes -- the narrow dilated
    // layers are bound by shared-memory bandwidth, not by the tensor pipe.  With the MMA spanning a CTA pair (M = 256) each CTA
    // reads its own A tile but only half of B (and fetches only half of every weight slot): 5 KB / 6 KB per MMA.  Taken for the
    // multi-tap layers with n-tiles <= 128 wide whose units can be grouped into quads of equal window phase (unit_index).
    // AID_TC2_CG2: 0 = never, 1 = wherever the shape allows, unset = where it measured faster with the whole kernel (B200, B = 8):
    // the 256-cout layers as two 128-wide n-tiles (0.499 -> 0.466 ms, 0.374 -> 0.351 ms) and the 96-channel layers (0.640 -> 0.619);
    // at 64 / 128 couts the main loop gains 3-11 % but the epilogue, which bounds those layers, runs slower next to the
    // pair's operand exchange (0.388 -> 0.402, 0.427 -> 0.437, 0.263 -> 0.280 ms).
    const int env_cg2 = getenv("AID_TC2_CG2") ? atoi(getenv("AID_TC2_CG2")) : -1;     // read per launch (tests switch it)
    static const int dbg = getenv("AID_TC_DEBUG") ? atoi(getenv("AID_TC_DEBUG")) : 0;
    p.cg2 = 0; p.qmode = 0;
    const bool cg2_wanted = env_cg2 > 0 || (env_cg2 < 0 && (p.n_ntiles == 2 || p.Ntile == 96));
    if (cg2_wanted && !(dbg & (2 | 32 | 8192)) && KT == 3 && p.ktb == 3 && p.Ntile <= 128 && p.Ntile % 16 == 0 && Cin % 16 == 0 && (Cin % 64 == 0 || Cin % 64 == 32) && num_sms % 2 == 0 &&
        num_sms >= 2 && !p.out_cl && !p.r_cl) {
        if (p.stream || p.tiles_t % 4 == 0) { p.cg2 = 1; p.qmode = 0; }
        else if (p.tiles_t == 2 && F % 8 == 0) { p.cg2 = 1; p.qmode = 1; }
        else if (p.tiles_t == 1 && F % 8 == 0) { p.cg2 = 1; p.qmode = 2; }
    }
    if (p.cg2) {
        p.n_pairs = (p.n_units + 3) / 4;       // quads
        p.n_tiles = p.n_pairs * p.n_ntiles;
        p.mg_pairs = div_magic(p.n_pairs);
    }
    p.b_slot_bytes = p.ktb * kt_bytes / (p.cg2 ? 2 : 1);
    p.nA = max(2, min(4, env_nA));
    p.a_slot_bytes = 2 * T2_ASLOT_UNIT;
    const int budget = 224 * 1024 - 1024 - T2_BAR_BYTES - T2_STAT_GATE_BYTES - T2_STAT_SMEM;
    while (p.nA > 2 && budget - p.nA * p.a_slot_bytes < 2 * p.b_slot_bytes) --p.nA;
    p.nB = min(8, (budget - p.nA * p.a_slot_bytes) / p.b_slot_bytes);
    if (p.nB < 2) throw CudaError(cudaErrorInvalidValue, "conv_tc2: shared memory budget", __FILE__, __LINE__);
    p.ncol_stride = p.Ntile <= 64 ? 64 : (p.Ntile <= 128 ? 128 : 256);
    p.acc_bufs = p.ncol_stride <= 128 ? 2 : 1;
    // epilogue warps: 16 (four column groups) when the n-tile splits into four groups of whole 8-column chunks (AID_TC2_EW=8 forces 8)
    // (16 epilogue warps with 16-column batches and 96 registers measured no faster than 8 -- AID_TC2_EW=16 keeps it reachable)
    const int env_ew = getenv("AID_TC2_EW") ? atoi(getenv("AID_TC2_EW")) : 8;
    const int ew = (env_ew == 16 && p.Ntile % 32 == 0) ? 16 : 8;
    // static epilogue for the shapes of the paper networks (AID_TC2_FASTEPI=0: generic one everywhere)
    const int env_fe = getenv("AID_TC2_FASTEPI") ? atoi(getenv("AID_TC2_FASTEPI")) : 1;
    p.epi_class = 0;
    if (env_fe && ew == 8 && !p.out_cl && !p.r_cl && !(dbg & (1 | 256 | 512 | 2048))) {
        const int hc = p.Ntile / 2, g8 = p.Ntot / 8;
        const bool st = ep.stats != nullptr;
        if (hc == 32 && (!st || g8 == 8)) p.epi_class = 1;
        else if (hc == 48 && (!st || g8 == 12)) p.epi_class = 2;
        else if (hc == 64 && (!st || g8 == 16)) p.epi_class = 3;
        else if (hc == 64 && g8 == 32) p.epi_class = 4;
        else if (hc == 128 && (!st || g8 == 32)) p.epi_class = 5;
    }
    // an epilogue warp owns Ntile / NCW columns: they must be whole statistics groups, at most 8 / NCW of them
    const int ncw = ew / 4, wcols = p.Ntile / ncw, gcn = p.Ntot / 8;
    if (ep.stats && p.n_ntiles != 1 && (wcols % gcn != 0 || wcols > (8 / ncw) * gcn))
        throw CudaError(cudaErrorInvalidValue, "conv_tc2: statistics groups do not align with the n-tiles", __FILE__, __LINE__);
    const size_t smem = 1024 + (size_t)p.nA * p.a_slot_bytes + (size_t)p.nB * p.b_slot_bytes + T2_BAR_BYTES + T2_STAT_GATE_BYTES + T2_STAT_SMEM;
    p.dbg = dbg;
    const int threads = (ew + 4) * 32;
    static SmemConfig cfg_8, cfg_16, cfg2_8, cfg2_16;
    if (p.cg2) {
        const int grid = 2 * min(p.n_tiles, num_sms / 2);     // __cluster_dims__(2, 1, 1)
        if (dbg & 4096) fprintf(stderr, "conv_tc2 cta_group::2: grid %d CTAs x %d threads, smem %zu, nA %d nB %d qmode %d\n", grid, threads, smem, p.nA, p.nB, p.qmode);
        if (ew == 16) { ensure_dyn_smem(conv_tc2_cg2_kernel<16>, smem, cfg2_16); conv_tc2_cg2_kernel<16><<<grid, threads, smem, s>>>(p); }
        else { ensure_dyn_smem(conv_tc2_cg2_kernel<8>, smem, cfg2_8); conv_tc2_cg2_kernel<8><<<grid, threads, smem, s>>>(p); }
    } else {
        const int grid = min(p.n_tiles, num_sms);
        if (ew == 16) { ensure_dyn_smem(conv_tc2_kernel<16>, smem, cfg_16); conv_tc2_kernel<16><<<grid, threads, smem, s>>>(p); }
        else { ensure_dyn_smem(conv_tc2_kernel<8>, smem, cfg_8); conv_tc2_kernel<8><<<grid, threads, smem, s>>>(p); }
    }
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
