// Fused dilated residual layer of conv_mode 2 for 64-channel (below) and 96-channel (second half of the file) blocks (unet.py:470-482):
// group-norm apply, adaLN modulation, GELU and the
// conversion to the fp16 tensor-core operand happen INSIDE the convolution kernel, so a layer reads its input once and writes
// its output once (8 B per element instead of 6 B for the operand pass + 10 B for conv_tc2_kernel).
//
// Generating the operand in the consumer is only affordable if every input row is transformed once although five output rows
// (the kf taps, dil rows apart) use it.  A work item is therefore a COMB: one clip, one 128-pixel t-tile and the rows
// f = r, r + dil, r + 2 dil, ... of one residue r mod dil.  Walking the comb, output row k needs the transformed rows k-2 .. k+2 of
// the same comb: they live in a ring of five shared-memory slots (130 pixels x 64 channels, fp16, K-major SWIZZLE_128B -- the
// layout conv_tc2 loads with a bulk copy is written here by the transform warps), each new output row costs ONE new input row.
// All 15 weight taps (64 x 64 fp16 each, 120 KB) stay resident in shared memory for the whole launch.
//   warps 0-7   epilogue (tc_epilogue.cuh: TMEM -> out = alpha * (acc * gate + x), statistics of the output for the next layer)
//   warps 8-15  transform: warp w owns channels 8w .. 8w+7 (one 16-byte operand chunk), lane l the pixels l, l+32, l+64, l+96 of
//               the tile; raw values are held one row ahead in registers, the row after that is pulled into L2 with evict_last
//               priority (the epilogue re-reads the row as the residual three output rows later: L2 hits)
//   warp 16     loads the weights once, then transforms the two halo pixels (t0 - 1, t0 + 128) of every row
//   warp 17     one elected thread issues the MMAs: per output row 5 kf x 3 kt x 4 k-steps of M = 128, N = 64, K = 16
// Ring protocol (global row sequence number n per CTA, slot = n % 5): row_ready[slot] <- the 8 transform warps + the halo warp;
// slot_free[slot] <- tcgen05.commit after the last MMAs that read the row (its kf = 0 use two output rows later, or the end of
// the comb).  With the kf taps issued in ascending order a row is needed last, at the very end of an output row, so the
// transform of row n+5 overlaps 1.6 output rows of MMAs.
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace aid {

static constexpr int CB_C = 64;                  // channels (in = out)
static constexpr int CB_EPI_WARPS = 8, CB_TR_WARPS = 8, CB_WARP_W = 16, CB_WARP_MMA = 17, CB_THREADS = 18 * 32;
static constexpr int CB_SLOT = 17408;            // one transformed row window: 136 rows x 128 B (130 used), 1 KB aligned
static constexpr int CB_RING = 5, CB_NACC = 4;
static constexpr int CB_W_BYTES = 15 * CB_C * 128;   // 15 taps x [64 couts][64 cins] fp16
static constexpr int CB_BAR_BYTES = 256, CB_GATE_BYTES = 4096, CB_STAT_BYTES = 16384;
static constexpr size_t CB_SMEM = 1024 + (size_t)CB_RING * CB_SLOT + CB_W_BYTES + CB_BAR_BYTES + CB_GATE_BYTES + CB_STAT_BYTES;

struct CombArgs {
    TV x, out;                       // layer input (operand source and residual) and output; must not overlap
    const double* stats_in; double n_per_group;   // (sum, sumsq) of x per (clip, group of 8 channels)
    const float* gamma; const float* affine; long long affine_bstride;
    const float* gate; long long gate_bstride;
    float alpha; double* stats_out;
    const __half* w;                 // launch_pack_weight_tc2 layout: [kf][kt][64 couts][64 cins], chunks swizzled by (cout & 7)
    int B, F, T, dil, tiles_t, n_items;
    int dbg;                         // AID_COMB_DEBUG=1 (tuning): block 0 prints the cycles its roles spent waiting
};

struct CombItem { int b, r, t0, K; };     // clip, comb residue, first pixel of the t-tile, rows of the comb
__device__ __forceinline__ CombItem comb_item(const CombArgs& p, int item) {
    CombItem it;
    const int tt = item % p.tiles_t, br = item / p.tiles_t;
    it.r = br % p.dil; it.b = br / p.dil; it.t0 = tt * 128;
    it.K = (p.F - it.r + p.dil - 1) / p.dil;
    return it;
}

// unit sequence of an epilogue thread: item -> its output rows, one accumulator each
struct CombUnitIter {
    const CombArgs& p; int pofs, step, acc_stride;
    int item, k = 0, ab = 0; uint32_t aph = 0;
    CombItem ci;
    __device__ CombUnitIter(const CombArgs& p_, int item0, int step_, int pofs_, int acc_stride_)
        : p(p_), pofs(pofs_), step(step_), acc_stride(acc_stride_), item(item0 - step_) { advance_item(); }
    __device__ __forceinline__ void advance_item() {      // next item with at least one row (dil > F leaves empty combs)
        do { item += step; if (item < p.n_items) ci = comb_item(p, item); } while (item < p.n_items && ci.K <= 0);
    }
    __device__ __forceinline__ bool next(EpiUnit& d) {
        if (item >= p.n_items) return false;
        const int f = ci.r + k * p.dil;
        d.b = ci.b; d.nt = 0; d.ok = true;
        d.pix = (long long)f * p.T + ci.t0 + pofs;
        d.tcol = (uint32_t)(ab * acc_stride); d.ab = ab; d.aph = aph; d.first = true; d.last = true;
        if (++ab == CB_NACC) { ab = 0; aph ^= 1; }
        if (++k == ci.K) { k = 0; advance_item(); }
        return true;
    }
};

__global__ void __launch_bounds__(CB_THREADS, 1) conv_comb_kernel(const __grid_constant__ CombArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem;
    uint8_t* wsm = smem + (size_t)CB_RING * CB_SLOT;
    uint8_t* bar_base = wsm + CB_W_BYTES;
    uint64_t* row_ready = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* slot_free = row_ready + CB_RING;
    uint64_t* tmem_full = slot_free + CB_RING;
    uint64_t* tmem_empty = tmem_full + CB_NACC;
    uint64_t* w_full = tmem_empty + CB_NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* gsm_base = reinterpret_cast<float*>(bar_base + CB_BAR_BYTES);
    double* sacc_base = reinterpret_cast<double*>(bar_base + CB_BAR_BYTES + CB_GATE_BYTES);
    const int item0 = blockIdx.x, istep = gridDim.x;

    if (warp == CB_WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < CB_RING; ++s) { mbar_init(row_ready + s, CB_TR_WARPS + 1); mbar_init(slot_free + s, 1); }
            for (int s = 0; s < CB_NACC; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, CB_EPI_WARPS); }
            mbar_init(w_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // per-channel operand scale of clip b, the arithmetic of gn_act_tc2_kernel: gamma * (1 + affine) / (unbiased std + eps)
    auto channel_scale = [&](int b, int c) {
        const double s1 = p.stats_in[((long long)b * 8 + (c >> 3)) * 2 + 0], s2 = p.stats_in[((long long)b * 8 + (c >> 3)) * 2 + 1];
        double var = (s2 - s1 * s1 / p.n_per_group) / (p.n_per_group - 1.0);
        var = var > 0.0 ? var : 0.0;
        const float inv = 1.f / ((float)sqrt(var) + 1e-7f);
        const float mod = p.affine ? (1.f + p.affine[b * p.affine_bstride + c]) : 1.f;
        return p.gamma[c] * mod * inv;
    };
    if (warp == CB_WARP_W) {
        // ===================== weights (resident for the whole launch), then the two halo pixels of every row =====================
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)CB_W_BYTES);
            for (int k = 0; k < 15; ++k) bulk_g2s(wsm + k * (CB_C * 128), p.w + (size_t)k * CB_C * 64, CB_C * 128, w_full);
        }
        __syncwarp();
        // lanes 0-7: window position 0 (pixel t0 - 1), lanes 8-15: position 129 (t0 + 128); lane & 7 = the 8-channel operand chunk.
        // One row of loads ahead of the arithmetic; zero outside the row.
        const int side = (lane >> 3) & 1, chunk = lane & 7;
        const bool act = lane < 16;
        const long long sc = p.x.sc;
        const uint32_t ring_u = smem_u32(ring);
        int n = 0, b_cur = -1;
        float cu[8], chh[8];
        for (int item = item0; item < p.n_items; item += istep) {
            const CombItem ci = comb_item(p, item);
            if (ci.b != b_cur) {
                b_cur = ci.b;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float sj = channel_scale(ci.b, 8 * chunk + j);
                    cu[j] = fabsf(sj) * 0.84932180028801904272f; chh[j] = 8.f * sj;
                }
            }
            const int th = side ? ci.t0 + 128 : ci.t0 - 1;
            const bool ld = act && th >= 0 && th < p.T;
            const float* xh = p.x.p + (long long)ci.b * p.x.sb + (long long)(8 * chunk) * sc + (ld ? th : 0);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (ld && ci.K > 0) ? __ldg(xh + (long long)ci.r * p.T + (long long)j * sc) : 0.f;
            for (int k = 0; k < ci.K; ++k, ++n) {
                float cv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) cv[j] = v[j];
                if (k + 1 < ci.K) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = ld ? __ldg(xh + (long long)(ci.r + (k + 1) * p.dil) * p.T + (long long)j * sc) : 0.f;
                }
                uint32_t hh[4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const float2 rh = gelu16_tc2_folded2(make_float2(cv[2 * pr], cv[2 * pr + 1]), make_float2(cu[2 * pr], cu[2 * pr + 1]), make_float2(chh[2 * pr], chh[2 * pr + 1]));
                    hh[pr] = pack_half2_sat(rh.x, rh.y);
                }
                const int slot = n % CB_RING;
                if (n >= CB_RING) mbar_wait(slot_free + slot, (uint32_t)((n / CB_RING) - 1) & 1u);
                if (act) {
                    const uint32_t row = side ? 129u : 0u;
                    const uint32_t addr = ring_u + (uint32_t)slot * CB_SLOT + row * 128u + (((uint32_t)chunk ^ (row & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hh[0]), "r"(hh[1]), "r"(hh[2]), "r"(hh[3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(row_ready + slot);
            }
        }
    } else if (warp >= CB_EPI_WARPS && warp < CB_EPI_WARPS + CB_TR_WARPS) {
        // ===================== transform: x row -> 16 * GELU(x * scale) fp16, K-major SWIZZLE_128B window =====================
        const int w = warp - CB_EPI_WARPS;             // channels 8w .. 8w+7 = operand chunk w
        int n = 0;                                     // global row sequence number of this CTA
        long long t_wait = 0, t_all = clock64();
        int b_cur = -1;
        // this warp's scale constants live in shared memory (the unused half of epilogue warp w's gate table), one LDS.128 per
        // channel pair: [pair][cu0, cu1, ch0, ch1] -- 16 registers the 96-register budget of an 18-warp CTA does not have
        float* sct = gsm_base + w * 128 + 64;
        const long long sc = p.x.sc;
        const uint32_t ring_u = smem_u32(ring);
        for (int item = item0; item < p.n_items; item += istep) {
            const CombItem ci = comb_item(p, item);
            if (ci.b != b_cur) {
                b_cur = ci.b;
                __syncwarp();
                if (lane < 8) {
                    const float sj = channel_scale(ci.b, 8 * w + lane);
                    sct[(lane >> 1) * 4 + (lane & 1)] = fabsf(sj) * 0.84932180028801904272f;
                    sct[(lane >> 1) * 4 + 2 + (lane & 1)] = 8.f * sj;
                }
                __syncwarp();
            }
            const float* xb = p.x.p + (long long)ci.b * p.x.sb + (long long)(8 * w) * sc + ci.t0 + lane;
            // The raw values of a row stay in registers one whole row ahead of the arithmetic: as soon as a channel pair of row k
            // has been consumed its registers receive the same pair of row k + 1 (L2 latency is ~800 cycles, a row of GELUs ~1500),
            // and the row after that is pulled into L2 meanwhile.
            float v[8][4];
            if (ci.K > 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[j][i] = __ldg(xb + (long long)ci.r * p.T + (long long)j * sc + i * 32);
            }
            for (int k = 0; k < ci.K; ++k, ++n) {
                const float* xr = xb + (long long)(ci.r + k * p.dil) * p.T;
                if (k + 2 < ci.K) {     // lane -> (channel lane / 4, 128-byte line lane % 4) of the row after the next
                    const float* nx = xr - lane + 2ll * p.dil * p.T + (long long)(lane >> 2) * sc + (lane & 3) * 32;
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nx));
                }
                const float* xn = xr + (long long)p.dil * p.T;
                const bool more = k + 1 < ci.K;
                uint32_t hp[4][4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sct + pr * 4);
                    const float2 cu2 = make_float2(s4.x, s4.y), ch2 = make_float2(s4.z, s4.w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r2 = gelu16_tc2_folded2(make_float2(v[2 * pr][i], v[2 * pr + 1][i]), cu2, ch2);
                        hp[i][pr] = pack_half2_sat(r2.x, r2.y);
                    }
                    if (more) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[2 * pr + c][i] = __ldg(xn + (long long)(2 * pr + c) * sc + i * 32);
                    }
                }
                const int slot = n % CB_RING;
                if (n >= CB_RING) { const long long t0 = clock64(); mbar_wait(slot_free + slot, (uint32_t)((n / CB_RING) - 1) & 1u); t_wait += clock64() - t0; }
                const uint32_t sbase = ring_u + (uint32_t)slot * CB_SLOT;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t row = 1u + (uint32_t)(i * 32 + lane);      // window position of pixel i * 32 + lane
                    const uint32_t addr = sbase + row * 128u + (((uint32_t)w ^ (row & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[i][0]), "r"(hp[i][1]), "r"(hp[i][2]), "r"(hp[i][3]) : "memory");
                }
                // generic-proxy writes -> visible to the tensor core (async proxy), then one arrival per warp
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(row_ready + slot);
            }
        }
        if (p.dbg && blockIdx.x == 0 && w == 0 && lane == 0) printf("comb transform: %d rows, total %lld cycles, waiting for a slot %lld\n", n, clock64() - t_all, t_wait);
    } else if (warp == CB_WARP_MMA) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(CB_C >> 3) << 17) | ((128u >> 4) << 24);   // F16 x F16 -> F32, K-major A/B, M = 128, N = 64
            const uint32_t adesc = desc_lo_sw128(smem_u32(ring)), bdesc = desc_lo_sw128(smem_u32(wsm));
            const uint32_t slot_d = CB_SLOT >> 4, tap_d = (CB_C * 128) >> 4;
            mbar_wait(w_full, 0);
            int n0 = 0;                  // sequence number of row 0 of the current item
            int ready = 0;               // rows [0, ready) of this CTA are known to be transformed
            int ab = 0; uint32_t aph = 0;
            long long t_te = 0, t_rr = 0, t_all = clock64();
            for (int item = item0; item < p.n_items; item += istep) {
                const CombItem ci = comb_item(p, item);
                for (int k = 0; k < ci.K; ++k) {
                    { const long long t0 = clock64(); mbar_wait(tmem_empty + ab, aph ^ 1); t_te += clock64() - t0; }
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(ab * CB_C);
                    uint32_t acc = 0u;
                    for (int kf = 0; kf < 5; ++kf) {
                        const int j = k + kf - 2;                 // comb row of this tap
                        if (j < 0 || j >= ci.K) continue;         // outside the plane: zero padding
                        const int n = n0 + j, slot = n % CB_RING;
                        if (n >= ready) { const long long t0 = clock64(); mbar_wait(row_ready + slot, (uint32_t)(n / CB_RING) & 1u); t_rr += clock64() - t0; ready = n + 1; tc_fence_after(); }
                        const uint32_t a = adesc + (uint32_t)slot * slot_d, b = bdesc + (uint32_t)(kf * 3) * tap_d;
                        tc_mma_k<1, 4>(d, a, b, idesc, acc);
                        tc_mma_k<1, 4>(d, a + 8u, b + tap_d, idesc, 1u);
                        tc_mma_k<1, 4>(d, a + 16u, b + 2u * tap_d, idesc, 1u);
                        acc = 1u;
                        // last reader of row j: its kf = 0 use (output row j + 2), or this output row if it is the last of the comb
                        if (kf == 0 || k == ci.K - 1) tc_commit(slot_free + slot);
                    }
                    tc_commit(tmem_full + ab);
                    if (++ab == CB_NACC) { ab = 0; aph ^= 1; }
                }
                // rows K-2, K-1 are read last by output row K-1 (released there); a comb shorter than 3 rows releases everything there too
                n0 += ci.K;
            }
            if (p.dbg && blockIdx.x == 0) printf("comb mma: %d rows, total %lld cycles, waiting for an accumulator %lld, for a row %lld\n", n0, clock64() - t_all, t_te, t_rr);
        }
        __syncwarp();
    } else if (warp < CB_EPI_WARPS) {
        EpiArgs ea{p.out, p.x, p.gate, p.gate_bstride, p.alpha, p.stats_out, CB_C, 1};
        CombUnitIter it(p, item0, istep, (warp & 3) * 32 + lane, CB_C);
        epilogue_fast<false, 16, 2, 8>(ea, it, warp, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CB_WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}


// ---- 96 channels -------------------------------------------------------------------------------------------------------------------
// The same scheme with K = 96 as a 64-channel group (128-byte rows, SWIZZLE_128B) plus a 32-channel group (64-byte rows,
// SWIZZLE_64B: the padded 128-channel layout of conv_tc2 would not leave room for the ring), and the weights (270 KB) streamed
// through a two-slot ring, one slot per (kf, channel group): three taps of [96 couts][64 cins] (36 KB), then three of [96 couts][32
// cins] (18 KB); two consecutive output rows of the comb share every slot.  MMA order per accumulator: kf, group, kt, k-step --
// the order of conv_tc2, so the fused layer reproduces the two-kernel path bit for bit and results do not depend on which of
// the two a batch size selects.
//   warps 0-7   epilogue (TMEM lane quadrant x column half, 4 batches of 12 columns = one statistics group each; 4 epilogue warps
//               measured twice as slow: the epilogue is a latency-bound stream per warp)
//   warps 8-13  transform, warp w = operand chunks w and w + 6 (two 8-channel chunks: 17 warps keep 96 registers per thread);
//               the raw values of the 8 channel pairs rotate through four register sets, loads four pairs ahead of the arithmetic
//   warp 14 halo pixels, 15 weight producer, 16 MMA issuer
static constexpr int C9 = 96, C9_EPI = 8, C9_TR = 6, C9_WARP_HALO = 14, C9_WARP_W = 15, C9_WARP_MMA = 16, C9_THREADS = 17 * 32;
static constexpr int C9_SLOT0 = 17408, C9_SLOT1 = 9216;      // group 0: 136 rows x 128 B; group 1: 144 rows x 64 B (130 used), 1 KB aligned
static constexpr int C9_WTAP0 = C9 * 128, C9_WTAP1 = C9 * 64;     // one tap of the 64- / 32-channel group: 12288 / 6144 bytes
static constexpr int C9_WG0 = 3 * C9_WTAP0, C9_WG1 = 3 * C9_WTAP1, C9_WKF = C9_WG0 + C9_WG1;   // per kf: 36864 + 18432 bytes
static constexpr int C9_WSLOT = C9_WG0;                            // ring slot (a group-1 stage fills half of it)
static constexpr int C9_NW = 2, C9_NACC = 4, C9_ACC_STRIDE = 128;
static constexpr size_t C9_SMEM = 1024 + (size_t)CB_RING * (C9_SLOT0 + C9_SLOT1) + (size_t)C9_NW * C9_WSLOT + CB_BAR_BYTES + CB_GATE_BYTES + CB_STAT_BYTES;

size_t comb_weight_halves(int C) { return C == C9 ? (size_t)5 * C9_WKF / 2 : (size_t)15 * CB_C * 64; }

// w[co][ci][5][3] fp32 -> per kf: [kt][96][64] halves (x 2^10), 16-byte chunks swizzled by (co & 7), then [kt][96][32] halves, chunks by ((co >> 1) & 3)
__global__ void pack_weight_comb96_kernel(const float* __restrict__ w, __half* __restrict__ wp) {
    const int total = 15 * C9 * C9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % C9, co = (i / C9) % C9, tap = i / (C9 * C9);     // tap = kf * 3 + kt
        const int kf = tap / 3, kt = tap % 3;
        float v = w[((long long)co * C9 + ci) * 15 + tap] * T2_W_SCALE;
        v = fminf(fmaxf(v, -60000.f), 60000.f);
        size_t o;
        if (ci < 64) o = (size_t)kf * (C9_WKF / 2) + (size_t)kt * (C9_WTAP0 / 2) + (size_t)co * 64 + ((((ci >> 3) ^ (co & 7)) << 3) | (ci & 7));
        else { const int c = ci - 64; o = (size_t)kf * (C9_WKF / 2) + C9_WG0 / 2 + (size_t)kt * (C9_WTAP1 / 2) + (size_t)co * 32 + ((((c >> 3) ^ ((co >> 1) & 3)) << 3) | (c & 7)); }
        wp[o] = __float2half_rn(v);
    }
}
void launch_pack_weight_comb(const float* w, __half* wp, int C, cudaStream_t s) {
    if (C != C9) throw CudaError(cudaErrorInvalidValue, "pack_weight_comb: 96 channels only (64 uses the conv_tc2 packing)", __FILE__, __LINE__);
    pack_weight_comb96_kernel<<<270, 512, 0, s>>>(w, wp);
    AID_COUNT_LAUNCH(1);
}

__global__ void __launch_bounds__(C9_THREADS, 1) conv_comb96_kernel(const __grid_constant__ CombArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring0 = smem;                                        // [5] group-0 windows
    uint8_t* ring1 = ring0 + (size_t)CB_RING * C9_SLOT0;          // [5] group-1 windows
    uint8_t* wring = ring1 + (size_t)CB_RING * C9_SLOT1;          // [4] weight slots
    uint8_t* bar_base = wring + (size_t)C9_NW * C9_WSLOT;
    uint64_t* row_ready = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* slot_free = row_ready + CB_RING;
    uint64_t* tmem_full = slot_free + CB_RING;
    uint64_t* tmem_empty = tmem_full + C9_NACC;
    uint64_t* w_full = tmem_empty + C9_NACC;
    uint64_t* w_empty = w_full + C9_NW;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + C9_NW);
    float* gsm_base = reinterpret_cast<float*>(bar_base + CB_BAR_BYTES);
    double* sacc_base = reinterpret_cast<double*>(bar_base + CB_BAR_BYTES + CB_GATE_BYTES);
    const int item0 = blockIdx.x, istep = gridDim.x;

    if (warp == C9_WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < CB_RING; ++s) { mbar_init(row_ready + s, C9_TR + 1); mbar_init(slot_free + s, 1); }
            for (int s = 0; s < C9_NACC; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, C9_EPI); }
            for (int s = 0; s < C9_NW; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // per-channel operand scale of clip b (gn_act_tc2_kernel's arithmetic); 12 channels per statistics group
    auto channel_scale = [&](int b, int c) {
        const int g = c / (C9 / 8);
        const double s1 = p.stats_in[((long long)b * 8 + g) * 2 + 0], s2 = p.stats_in[((long long)b * 8 + g) * 2 + 1];
        double var = (s2 - s1 * s1 / p.n_per_group) / (p.n_per_group - 1.0);
        var = var > 0.0 ? var : 0.0;
        const float inv = 1.f / ((float)sqrt(var) + 1e-7f);
        const float mod = p.affine ? (1.f + p.affine[b * p.affine_bstride + c]) : 1.f;
        return p.gamma[c] * mod * inv;
    };
    // shared-memory address of the 16-byte chunk `chunk` (8 channels) of window row `row` in ring slot `slot`
    auto chunk_addr = [&](int slot, uint32_t row, int chunk) -> uint32_t {
        if (chunk < 8) return smem_u32(ring0) + (uint32_t)slot * C9_SLOT0 + row * 128u + (((uint32_t)chunk ^ (row & 7u)) << 4);
        return smem_u32(ring1) + (uint32_t)slot * C9_SLOT1 + row * 64u + ((((uint32_t)chunk - 8u) ^ ((row >> 1) & 3u)) << 4);
    };

    if (warp == C9_WARP_W) {
        // ===================== weight producer: one slot per (kf, channel group) of every PAIR of output rows =====================
        if (lane == 0) {
            int ws = 0; uint32_t wph = 0;
            for (int item = item0; item < p.n_items; item += istep) {
                const CombItem ci = comb_item(p, item);
                for (int k0 = 0; k0 < ci.K; k0 += 2)
                    for (int kf = 0; kf < 5; ++kf) {
                        const int j0 = k0 + kf - 2, j1 = j0 + 1;       // input rows of the taps of output rows k0 and k0 + 1
                        const bool v0 = j0 >= 0 && j0 < ci.K, v1 = k0 + 1 < ci.K && j1 >= 0 && j1 < ci.K;
                        if (!(v0 || v1)) continue;
                        for (int g = 0; g < 2; ++g) {
                            const uint32_t bytes = g ? (uint32_t)C9_WG1 : (uint32_t)C9_WG0;
                            mbar_wait(w_empty + ws, wph ^ 1);
                            mbar_expect_tx(w_full + ws, bytes);
                            bulk_g2s(wring + (size_t)ws * C9_WSLOT, reinterpret_cast<const uint8_t*>(p.w) + (size_t)kf * C9_WKF + (g ? C9_WG0 : 0), bytes, w_full + ws);
                            if (++ws == C9_NW) { ws = 0; wph ^= 1; }
                        }
                    }
            }
        }
        __syncwarp();
    } else if (warp == C9_WARP_HALO) {
        // ===================== the two halo pixels of every row: lane -> (side = lane / 12, chunk = lane % 12) =====================
        const int side = lane / 12, chunk = lane % 12;
        const bool act = lane < 24;
        const long long sc = p.x.sc;
        int n = 0, b_cur = -1;
        float cu[8], chh[8];
        for (int item = item0; item < p.n_items; item += istep) {
            const CombItem ci = comb_item(p, item);
            if (ci.b != b_cur) {
                b_cur = ci.b;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float sj = act ? channel_scale(ci.b, 8 * chunk + j) : 0.f;
                    cu[j] = fabsf(sj) * 0.84932180028801904272f; chh[j] = 8.f * sj;
                }
            }
            const int th = side ? ci.t0 + 128 : ci.t0 - 1;
            const bool ld = act && th >= 0 && th < p.T;
            const float* xh = p.x.p + (long long)ci.b * p.x.sb + (long long)(act ? 8 * chunk : 0) * sc + (ld ? th : 0);
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (ld && ci.K > 0) ? __ldg(xh + (long long)ci.r * p.T + (long long)j * sc) : 0.f;
            for (int k = 0; k < ci.K; ++k, ++n) {
                float cv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) cv[j] = v[j];
                if (k + 1 < ci.K) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = ld ? __ldg(xh + (long long)(ci.r + (k + 1) * p.dil) * p.T + (long long)j * sc) : 0.f;
                }
                uint32_t hh[4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const float2 rh = gelu16_tc2_folded2(make_float2(cv[2 * pr], cv[2 * pr + 1]), make_float2(cu[2 * pr], cu[2 * pr + 1]), make_float2(chh[2 * pr], chh[2 * pr + 1]));
                    hh[pr] = pack_half2_sat(rh.x, rh.y);
                }
                const int slot = n % CB_RING;
                if (n >= CB_RING) mbar_wait(slot_free + slot, (uint32_t)((n / CB_RING) - 1) & 1u);
                if (act) {
                    const uint32_t addr = chunk_addr(slot, side ? 129u : 0u, chunk);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hh[0]), "r"(hh[1]), "r"(hh[2]), "r"(hh[3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(row_ready + slot);
            }
        }
    } else if (warp >= C9_EPI && warp < C9_EPI + C9_TR) {
        // ===================== transform: warp w = chunks w, w + 6; lane l = pixels l, l+32, l+64, l+96 =====================
        const int w = warp - C9_EPI;
        int n = 0, b_cur = -1;
        float* sct = gsm_base + w * 128 + 64;          // [8 pairs][cu0, cu1, ch0, ch1] in the unused half of epilogue warp w's gate table
        const long long sc = p.x.sc;
        for (int item = item0; item < p.n_items; item += istep) {
            const CombItem ci = comb_item(p, item);
            if (ci.b != b_cur) {
                b_cur = ci.b;
                __syncwarp();
                if (lane < 16) {
                    const int c = 8 * (w + 6 * (lane >> 3)) + (lane & 7);
                    const float sj = channel_scale(ci.b, c);
                    sct[(lane >> 1) * 4 + (lane & 1)] = fabsf(sj) * 0.84932180028801904272f;
                    sct[(lane >> 1) * 4 + 2 + (lane & 1)] = 8.f * sj;
                }
                __syncwarp();
            }
            // channel 8 (w + 6 h) + 2 q + c of pair position s = 4 h + q lives at xb + pair_off(s) + c * sc
            const float* xb = p.x.p + (long long)ci.b * p.x.sb + ci.t0 + lane;
            auto pair_ptr = [&](int srow, int s) { return xb + (long long)srow * p.T + (long long)(8 * (w + 6 * (s >> 2)) + 2 * (s & 3)) * sc; };
            float v[4][2][4];       // register set s & 3 holds pair position s; refilled with position s + 4 as soon as it is consumed
            if (ci.K > 0) {
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const float* q = pair_ptr(ci.r, s4);
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[s4][c][i] = __ldg(q + (long long)c * sc + i * 32);
                }
            }
            for (int k = 0; k < ci.K; ++k, ++n) {
                const int frow = ci.r + k * p.dil;
                if (k + 2 < ci.K && lane < 16) {     // L2 prefetch of the row after the next: lane -> (chunk half lane / 8 ... 16 channels x 4 lines = 64 lines, 4 per lane)
                    const float* nx = p.x.p + (long long)ci.b * p.x.sb + (long long)(frow + 2 * p.dil) * p.T + ci.t0 + (long long)(8 * (w + 6 * (lane >> 3)) + (lane & 7)) * sc;
#pragma unroll
                    for (int l4 = 0; l4 < 4; ++l4) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nx + l4 * 32));
                }
                const bool more = k + 1 < ci.K;
                const int slot = n % CB_RING;
                uint32_t hp[4][4];
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    const float4 s4 = *reinterpret_cast<const float4*>(sct + s8 * 4);
                    const float2 cu2 = make_float2(s4.x, s4.y), ch2 = make_float2(s4.z, s4.w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 r2 = gelu16_tc2_folded2(make_float2(v[s8 & 3][0][i], v[s8 & 3][1][i]), cu2, ch2);
                        hp[i][s8 & 3] = pack_half2_sat(r2.x, r2.y);
                    }
                    if (s8 < 4 || more) {       // position s8 + 4: the other chunk of this row, or the first chunk of the next row
                        const float* q = s8 < 4 ? pair_ptr(frow, s8 + 4) : pair_ptr(frow + p.dil, s8 - 4);
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[s8 & 3][c][i] = __ldg(q + (long long)c * sc + i * 32);
                    }
                    if ((s8 & 3) == 3) {        // a chunk is complete: four 16-byte stores per thread
                        if (s8 == 3 && n >= CB_RING) mbar_wait(slot_free + slot, (uint32_t)((n / CB_RING) - 1) & 1u);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t addr = chunk_addr(slot, 1u + (uint32_t)(i * 32 + lane), w + 6 * (s8 >> 2));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[i][0]), "r"(hp[i][1]), "r"(hp[i][2]), "r"(hp[i][3]) : "memory");
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(row_ready + slot);
            }
        }
    } else if (warp == C9_WARP_MMA) {
        // ===================== MMA issuer: 5 kf x 3 kt x (4 + 2) k-steps of M = 128, N = 96, K = 16 per output row =====================
        // Output rows are issued in PAIRS (k0, k0 + 1) that share every weight slot (the stream of 270 KB per row through a 72 KB ring
        // bounded the kernel: the issuer waited for weights a quarter of its time): per (kf, kt) slot six MMAs for row k0 (input row
        // k0 + kf - 2) and six for row k0 + 1 (input row k0 + kf - 1).  The pair needs input rows k0 - 2 .. k0 + 3 -- one more than the
        // ring holds: row k0 - 2 is read last by (k0, kf = 0) and row k0 - 1 by (k0, kf = 1), their slots are released right there, and
        // row k0 + 3, first needed at kf = 4, is transformed into the freed slot meanwhile.
        if (elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(C9 >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a0desc = desc_lo_sw128(smem_u32(ring0)), a1desc = desc_lo_sw128(smem_u32(ring1)), wdesc = desc_lo_sw128(smem_u32(wring));
            int n0 = 0, ready = 0, ws = 0, g = 0; uint32_t wph = 0;      // g: output rows issued so far (accumulator g % NACC, phase (g / NACC) & 1)
            long long t_te = 0, t_rr = 0, t_w = 0, t_all = clock64();
            auto wait_row = [&](int n) {
                if (n >= ready) { const long long t0 = clock64(); mbar_wait(row_ready + n % CB_RING, (uint32_t)(n / CB_RING) & 1u); t_rr += clock64() - t0; ready = n + 1; tc_fence_after(); }
            };
            for (int item = item0; item < p.n_items; item += istep) {
                const CombItem ci = comb_item(p, item);
                for (int k0 = 0; k0 < ci.K; k0 += 2) {
                    const bool two = k0 + 1 < ci.K;
                    const int ab0 = g % C9_NACC, ab1 = (g + 1) % C9_NACC;
                    { const long long t0 = clock64();
                      mbar_wait(tmem_empty + ab0, (uint32_t)((g / C9_NACC) & 1) ^ 1u);
                      if (two) mbar_wait(tmem_empty + ab1, (uint32_t)(((g + 1) / C9_NACC) & 1) ^ 1u);
                      t_te += clock64() - t0; }
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + (uint32_t)(ab0 * C9_ACC_STRIDE), d1 = tmem_base + (uint32_t)(ab1 * C9_ACC_STRIDE);
                    uint32_t acc0 = 0u, acc1 = 0u;
                    for (int kf = 0; kf < 5; ++kf) {
                        const int j0 = k0 + kf - 2, j1 = j0 + 1;
                        const bool v0 = j0 >= 0 && j0 < ci.K, v1 = two && j1 >= 0 && j1 < ci.K;
                        if (!(v0 || v1)) continue;
                        const int s0 = (n0 + j0) % CB_RING, s1 = (n0 + j1) % CB_RING;      // (unused when the row is out of range)
                        if (v0) wait_row(n0 + j0);
                        if (v1) wait_row(n0 + j1);
                        const uint32_t a00 = a0desc + (uint32_t)s0 * (C9_SLOT0 >> 4), a01 = a1desc + (uint32_t)s0 * (C9_SLOT1 >> 4);
                        const uint32_t a10 = a0desc + (uint32_t)s1 * (C9_SLOT0 >> 4), a11 = a1desc + (uint32_t)s1 * (C9_SLOT1 >> 4);
                        for (int g = 0; g < 2; ++g) {        // 64-channel group (four k-steps per tap), then the 32-channel group (two)
                            { const long long t0 = clock64(); mbar_wait(w_full + ws, wph); t_w += clock64() - t0; }
                            tc_fence_after();
                            const uint32_t b = wdesc + (uint32_t)ws * (C9_WSLOT >> 4);
#pragma unroll
                            for (int kt = 0; kt < 3; ++kt) {        // tap kt: one pixel row (128 / 64 bytes) further in the window
                                if (g == 0) {
                                    if (v0) { tc_mma_k<1, 4>(d0, a00 + (uint32_t)kt * 8u, b + (uint32_t)kt * (C9_WTAP0 >> 4), idesc, acc0); acc0 = 1u; }
                                    if (v1) { tc_mma_k<1, 4>(d1, a10 + (uint32_t)kt * 8u, b + (uint32_t)kt * (C9_WTAP0 >> 4), idesc, acc1); acc1 = 1u; }
                                } else {
                                    if (v0) tc_mma_k2_sw64(d0, a01 + (uint32_t)kt * 4u, b + (uint32_t)kt * (C9_WTAP1 >> 4), idesc, 1u);
                                    if (v1) tc_mma_k2_sw64(d1, a11 + (uint32_t)kt * 4u, b + (uint32_t)kt * (C9_WTAP1 >> 4), idesc, 1u);
                                }
                            }
                            tc_commit(w_empty + ws);
                            if (++ws == C9_NW) { ws = 0; wph ^= 1; }
                        }
                        // last readers inside a pair: input row k0 - 2 at (k0, kf = 0), input row k0 - 1 at (k0, kf = 1)
                        if (kf <= 1 && v0) tc_commit(slot_free + s0);
                    }
                    tc_commit(tmem_full + ab0);
                    if (two) tc_commit(tmem_full + ab1);
                    g += two ? 2 : 1;
                    if (k0 + 2 >= ci.K)         // last pair of the comb: rows k0 .. K-1 have no later reader
                        for (int j = k0; j < ci.K; ++j) tc_commit(slot_free + (n0 + j) % CB_RING);
                }
                n0 += ci.K;
            }
            if (p.dbg && blockIdx.x == 0)
                printf("comb96 mma: %d rows, total %lld cycles, waiting for an accumulator %lld, for a row %lld, for weights %lld\n", n0, clock64() - t_all, t_te, t_rr, t_w);
        }
        __syncwarp();
    } else if (warp < C9_EPI) {
        EpiArgs ea{p.out, p.x, p.gate, p.gate_bstride, p.alpha, p.stats_out, C9, 1};
        CombUnitIter it(p, item0, istep, (warp & 3) * 32 + lane, C9_ACC_STRIDE);
        epilogue_fast<false, 12, 4, 12, CombUnitIter, C9_EPI>(ea, it, warp, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == C9_WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool conv_comb_supported(int C, int F, int T, int dil) { return (C == CB_C || C == C9) && T % 128 == 0 && dil >= 1 && F >= 1; }
// A comb is one CTA's serial work item: with few clips there are fewer combs (B * dil * T / 128) than SMs and the un-fused path, whose
// tiles are single rows, is faster (measured at batch 1: 18.4 ms per forward with the fused layers, 12.9 ms without).
bool conv_comb_worthwhile(int B, int T, int dil, int num_sms) { return (long long)B * dil * (T / 128) * 4 >= 3ll * num_sms; }

// out = alpha * (x + gate * conv5x3_dil(GELU(GroupNorm(x) * (1 + affine)))), statistics of out -> stats_out (may be null)
void launch_conv_comb(const TV& x, const double* stats_in, long long n_per_group, const float* gamma, const float* affine, long long affine_bstride,
                      const __half* wp, int dil, const TV& out, const ConvEpilogue& ep, int num_sms, cudaStream_t s) {
    if (!conv_comb_supported(x.C, x.F, x.T, dil) || out.C != x.C) throw CudaError(cudaErrorInvalidValue, "conv_comb: unsupported shape", __FILE__, __LINE__);
    if (ep.R.p != x.p || ep.R2.p) throw CudaError(cudaErrorInvalidValue, "conv_comb: the residual must be the layer input", __FILE__, __LINE__);
    if (out.p == x.p) throw CudaError(cudaErrorInvalidValue, "conv_comb: in-place update is not possible (t-tile halos)", __FILE__, __LINE__);
    CombArgs p{};
    p.x = x; p.out = out; p.stats_in = stats_in; p.n_per_group = (double)n_per_group; p.gamma = gamma; p.affine = affine; p.affine_bstride = affine_bstride;
    p.gate = ep.gate; p.gate_bstride = ep.gate_bstride; p.alpha = ep.alpha; p.stats_out = ep.stats; p.w = wp;
    p.B = x.B; p.F = x.F; p.T = x.T; p.dil = dil; p.tiles_t = x.T / 128;
    p.n_items = x.B * dil * p.tiles_t;
    static const int dbg = getenv("AID_COMB_DEBUG") ? atoi(getenv("AID_COMB_DEBUG")) : 0;
    p.dbg = dbg;
    static SmemConfig configured, configured96;
    if (x.C == C9) {     // wp: launch_pack_weight_comb layout
        ensure_dyn_smem(conv_comb96_kernel, C9_SMEM, configured96);
        conv_comb96_kernel<<<std::min(p.n_items, num_sms), C9_THREADS, C9_SMEM, s>>>(p);
    } else {             // wp: launch_pack_weight_tc2 layout
        ensure_dyn_smem(conv_comb_kernel, CB_SMEM, configured);
        conv_comb_kernel<<<std::min(p.n_items, num_sms), CB_THREADS, CB_SMEM, s>>>(p);
    }
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
