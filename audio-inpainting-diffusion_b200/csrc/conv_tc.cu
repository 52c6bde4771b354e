// tcgen05 (5th-gen tensor core) implicit-GEMM kernel for the network's convolutions with the gate / residual /
// statistics epilogue fused:
//   * the dilated 5x3 residual-layer convolutions (unet.py:433-436, 482) -- 95 % of the forward's FLOPs,
//   * the 1x1 "H" convolutions of the init / out blocks (unet.py:675, 690, 719), the block projections
//     proj_in / res_conv (unet.py:412-415) and the attention qk Conv1d (unet.py:321, 355), all as taps = 1.
//
// Precision: error-compensated split fp16.  Activations are written as a = a_hi + a_lo (two fp16, scaled by 2^4),
// weights are pre-split the same way (scaled by 2^10); the kernel issues a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
// (kind::f16, fp32 accumulation in TMEM).  The dropped a_lo*w_lo term is 2^-22 relative, so the result is fp32-grade
// (the 1e-3 parity bar rules out single bf16/tf32 products through 75 stacked residual layers).
//
// GEMM view per "unit" = 128 consecutive pixels of one (clip, frequency row):
//     D[128 px, Ntile couts] += A[128 px, 16 ch] * B[Ntile, 16 ch]^T     for every (kf, kt, 16-channel step)
// Layouts (all K-major, SWIZZLE_NONE canonical: core matrix = 8 rows x 16 B, SBO = 128 B, LBO = plane stride):
//   activations in HBM : [B][C/8][F][T+2][8] fp16 (hi and lo arrays); one zero pixel each side of T = the conv's zero
//                        padding along T.  A 130-pixel window of one row is ONE contiguous 2080-byte run, fetched with
//                        cp.async.bulk; the kt taps are the same window with the descriptor start advanced by 16 B.
//   weights in HBM     : [n-tile][kf][Cin/16][hi|lo][kt][2 chunks][Ntile][8] fp16 -> one contiguous bulk copy per stage.
//   zero padding along F: a tap row outside [0,F) is skipped (no MMA issued), never loaded.
// One CTA tile = 2 units x one n-tile (two accumulators in TMEM share every weight stage), persistent over tiles;
// warp 0 = bulk-copy producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = epilogue (TMEM -> registers -> NCHW).
#include <cuda_fp16.h>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace aid {

static constexpr int TC_THREADS = 576;        // producer warp, MMA warp, 16 epilogue warps
static constexpr int TC_PLANE = 130 * 16;     // bytes of one (16 B chunk) x (130 pixel) plane of A in smem
static constexpr int TC_MAX_KPS = 4;          // most 16-channel k-steps per pipeline stage
static constexpr float TC_A_SCALE = 16.f, TC_W_SCALE = 1024.f, TC_OUT_SCALE = 1.f / (16.f * 1024.f);

struct TcConvArgs {
    const __half* a_hi; const __half* a_lo; const __half* w;
    TV out, R;
    const float* gate; long long gate_bstride;
    float alpha; double* stats;
    int B, Cin, Ntot, Ntile, n_ntiles, F, T, Tp, dil;
    int PF, rows_total, stream, units_per_b;   // zero pad rows above/below each plane; stream mode: units tile the padded pixel stream
    int KF, KT, kt_shift, kps;                 // taps along F / T, first tap's pixel shift inside the window, k-steps per stage
    int parts;                                 // 2: split fp16 (hi, lo) operands, 3 MMAs per tap;  1: single fp16 operands, 1 MMA per tap
    int tiles_t, n_units, n_pairs, n_tiles, nstages, acc_bufs, ncol_stride;
    int b_kstep_bytes, a_bytes, stage_bytes;
    int dbg;  // AID_TC_DEBUG bits (tuning only): 1 skip epilogue body, 2 skip MMAs, 4 skip A loads, 8 skip B loads
};

// A unit = 128 consecutive output positions of one clip.
//   row mode    (T % 128 == 0): 128 pixels of one frequency row (PF = 0, taps outside [0,F) are skipped exactly);
//   stream mode (otherwise)   : 128 consecutive positions of the padded pixel stream [F rows][T+2] -- rows shorter than
//                               128 pixels share a tile (the pad pixels between rows are the zero padding along T, their
//                               outputs are discarded); PF zero rows above/below make every tap address valid.
struct UnitInfo { int exists, b, f_lo, f_hi, win_start, o0, seg_px; };

__device__ __forceinline__ UnitInfo unit_info(const TcConvArgs& p, int u) {
    UnitInfo i;
    i.exists = u < p.n_units;
    if (p.stream) {
        const int k = u % p.units_per_b;
        i.b = u / p.units_per_b;
        i.o0 = k * 128;                               // first output position, relative to row 0 of the real rows
        i.f_lo = i.o0 / p.Tp;
        i.f_hi = min(p.F - 1, (i.o0 + 127) / p.Tp);
        i.win_start = p.PF * p.Tp + i.o0 - 1;         // window = positions [o0-1, o0+129) of the padded plane
        i.seg_px = 130;
    } else {
        const int tt = u % p.tiles_t, r = u / p.tiles_t;
        const int f = r % p.F, t0 = tt * 128;
        i.b = r / p.F; i.f_lo = i.f_hi = f;
        i.o0 = f * p.Tp + t0 + 1;
        i.win_start = f * p.Tp + t0;
        i.seg_px = min(130, p.Tp - t0);
    }
    return i;
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(TcConvArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* bar_base = smem + (size_t)p.nstages * p.stage_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* empty = full + 12;
    uint64_t* tmem_full = empty + 12;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < p.nstages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 512); }
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

    const int KS = p.Cin >> 4;
    const int c8_total = p.Cin >> 3;
    const int b_bytes_full = p.b_kstep_bytes * p.kps;

    if (warp == 0) {
        // ===================== producer: bulk copies HBM/L2 -> shared (whole warp loops; lane 0: weights, lanes 1..: A) =====================
        int stage = 0; uint32_t phase = 0;
        // lane l >= 1: A copy of unit ai, k-step akk of the stage, (hi|lo), 16-byte chunk ac
        const int aidx = lane - 1;
        const int per_unit = p.kps * p.parts * 2;
        const int ai = aidx / per_unit, arem = aidx % per_unit;
        const int akk = arem / (p.parts * 2), ahl = (arem / 2) % p.parts, ac = arem & 1;
        const bool a_lane = lane >= 1 && aidx < 2 * per_unit;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int pair = tile % p.n_pairs, nt = tile / p.n_pairs;
            const UnitInfo u0 = unit_info(p, 2 * pair), u1 = unit_info(p, 2 * pair + 1);
            for (int kf = 0; kf < p.KF; ++kf) {
                const int foff = (kf - p.KF / 2) * p.dil;
                const bool v0 = u0.exists && u0.f_hi + foff >= 0 && u0.f_lo + foff < p.F;
                const bool v1 = u1.exists && u1.f_hi + foff >= 0 && u1.f_lo + foff < p.F;
                if (!(v0 || v1)) continue;
                for (int ks0 = 0; ks0 < KS; ks0 += p.kps) {
                    const int nk = min(p.kps, KS - ks0);
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sb = smem + (size_t)stage * p.stage_bytes;
                    if (lane == 0) {
                        const uint32_t bb = (uint32_t)(p.b_kstep_bytes * nk);
                        uint32_t bytes = (p.dbg & 8) ? 0u : bb;
                        if (!(p.dbg & 4)) bytes += (uint32_t)(nk * p.parts) * ((v0 ? 2u * u0.seg_px * 16u : 0u) + (v1 ? 2u * u1.seg_px * 16u : 0u));
                        mbar_expect_tx(full + stage, bytes);
                        if (!(p.dbg & 8))
                            bulk_g2s(sb, p.w + ((size_t)(nt * p.KF + kf) * KS + ks0) * (p.b_kstep_bytes >> 1), bb, full + stage);
                    }
                    __syncwarp();
                    if (a_lane && akk < nk && !(p.dbg & 4)) {
                        const UnitInfo& u = ai ? u1 : u0;
                        if (ai ? v1 : v0) {
                            const size_t off = (((size_t)u.b * c8_total + (2 * (ks0 + akk) + ac)) * p.rows_total * p.Tp +
                                                (size_t)(u.win_start + foff * p.Tp)) * 8;
                            bulk_g2s(sb + b_bytes_full + (((ai * p.kps + akk) * p.parts + ahl) * 2 + ac) * TC_PLANE, (ahl ? p.a_lo : p.a_hi) + off,
                                     (uint32_t)u.seg_px * 16u, full + stage);
                        }
                    }
                    if (++stage == p.nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, lane 0 issues) =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Ntile >> 3) << 17) | ((128u >> 4) << 24);  // F16 x F16 -> F32, K-major A/B
        const uint32_t b_lbo = (uint32_t)p.Ntile * 16u;
        int stage = 0; uint32_t phase = 0; int ab = 0; uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int pair = tile % p.n_pairs;
            const UnitInfo u0 = unit_info(p, 2 * pair), u1 = unit_info(p, 2 * pair + 1);
            mbar_wait(tmem_empty + ab, aphase ^ 1);
            tc_fence_after();
            uint32_t started[2] = {0u, 0u};
            for (int kf = 0; kf < p.KF; ++kf) {
                const int foff = (kf - p.KF / 2) * p.dil;
                const bool v0 = u0.exists && u0.f_hi + foff >= 0 && u0.f_lo + foff < p.F;
                const bool v1 = u1.exists && u1.f_hi + foff >= 0 && u1.f_lo + foff < p.F;
                if (!(v0 || v1)) continue;
                for (int ks0 = 0; ks0 < KS; ks0 += p.kps) {
                    const int nk = min(p.kps, KS - ks0);
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sb = smem_u32(smem + (size_t)stage * p.stage_bytes);
                        const uint32_t sa = sb + (uint32_t)b_bytes_full;
                        if (!(p.dbg & 2)) {
                            for (int i = 0; i < 2; ++i) {
                                if (!(i ? v1 : v0)) continue;
                                const uint32_t d = tmem_base + (uint32_t)(ab * 2 * p.ncol_stride + i * p.ncol_stride);
                                for (int kk = 0; kk < nk; ++kk) {
                                    const uint32_t a0 = sa + (uint32_t)(((i * p.kps + kk) * p.parts) * 2) * TC_PLANE;
                                    const uint32_t b0 = sb + (uint32_t)(kk * p.b_kstep_bytes);
                                    for (int kt = 0; kt < p.KT; ++kt) {
                                        const uint32_t sh = (uint32_t)(kt + p.kt_shift) * 16u;
                                        const uint64_t a_hi = make_desc(a0 + sh, TC_PLANE, 128);
                                        const uint64_t b_hi = make_desc(b0 + (uint32_t)((0 * p.KT + kt) * 2) * b_lbo, b_lbo, 128);
                                        tc_mma_f16(d, a_hi, b_hi, idesc, started[i]);
                                        started[i] = 1u;
                                        if (p.parts == 2) {
                                            const uint64_t a_lo = make_desc(a0 + 2 * TC_PLANE + sh, TC_PLANE, 128);
                                            const uint64_t b_lo = make_desc(b0 + (uint32_t)((1 * p.KT + kt) * 2) * b_lbo, b_lbo, 128);
                                            tc_mma_f16(d, a_lo, b_hi, idesc, 1u);
                                            tc_mma_f16(d, a_hi, b_lo, idesc, 1u);
                                        }
                                    }
                                }
                            }
                        }
                        tc_commit(empty + stage);  // frees the smem slot when these MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == p.nstages) { stage = 0; phase ^= 1; }
                }
            }
            if (lane == 0) tc_commit(tmem_full + ab);  // accumulators of this tile are complete
            __syncwarp();
            if (++ab == p.acc_bufs) { ab = 0; aphase ^= 1; }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> out = alpha*(acc*gate + R), statistics =====================
        // 16 warps: warp e owns TMEM lane quadrant (warpid & 3) and a contiguous column range (quarter, or half when the
        // tile width is not a multiple of 32).  The loop body is kept short on purpose: the epilogue is instruction/latency
        // bound, not bandwidth bound (few resident warps), so every per-column instruction counts.
        const int e = warp - 2;
        const int q = warp & 3;
        const int cw = e >> 2;
        const int ncw = (p.Ntile & 31) ? 2 : 4;
        const int ncols = p.Ntile / ncw;  // multiple of 8
        const bool active = cw < ncw;
        const int cbeg = cw * ncols;
        int ab = 0; uint32_t aphase = 0;
        const int gcn = p.Ntot / 8;
        const float al = p.alpha, gs = TC_OUT_SCALE * p.alpha;
        const long long osc = p.out.sc, rsc = p.R.sc;
        double* sst = reinterpret_cast<double*>(bar_base + 256) + e * 16;  // this warp's (group, {sum, sumsq}) accumulators
        if (lane < 16) sst[lane] = 0.0;
        __syncwarp();
        int b_cur = -1;
        auto flush_global = [&]() {
            __syncwarp();  // lane 0's accumulator updates (flush_group) must be visible to lanes 0..15
            if (p.stats && b_cur >= 0 && lane < 16) {
                const double v = sst[lane];
                if (v != 0.0) atomicAdd(p.stats + (long long)b_cur * 16 + lane, v);
                sst[lane] = 0.0;
            }
            __syncwarp();
        };
        auto flush_group = [&](float s, float qq, int g) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); qq += __shfl_xor_sync(0xffffffffu, qq, o); }
            if (lane == 0) { sst[g * 2 + 0] += (double)s; sst[g * 2 + 1] += (double)qq; }
        };
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int pair = tile % p.n_pairs, nt = tile / p.n_pairs;
            const int co_base = nt * p.Ntile;
            mbar_wait(tmem_full + ab, aphase);
            tc_fence_after();
#pragma unroll 1
            for (int i = 0; i < 2 && active && !(p.dbg & 1); ++i) {
                const UnitInfo u = unit_info(p, 2 * pair + i);
                if (!u.exists) continue;
                if (u.b != b_cur) { flush_global(); b_cur = u.b; }
                const int o = u.o0 + q * 32 + lane;           // output position in the padded stream of the real rows
                const int row = o / p.Tp, tp = o - row * p.Tp;
                const bool ok = tp >= 1 && tp <= p.T && row <= u.f_hi;
                const long long pix = (long long)row * p.T + (tp - 1);
                float* po = p.out.p + (long long)u.b * p.out.sb + (long long)(co_base + cbeg) * osc + pix;
                const float* pr = p.R.p + (long long)u.b * p.R.sb + (long long)(co_base + cbeg) * rsc + pix;
                const bool hasr = ok && p.R.p != nullptr;
                const float* gate = p.gate ? p.gate + (long long)u.b * p.gate_bstride + co_base + cbeg : nullptr;
                const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 2 * p.ncol_stride + i * p.ncol_stride + cbeg);
                int grp = (co_base + cbeg) / gcn;
                int left = (grp + 1) * gcn - (co_base + cbeg);  // columns left in the current statistics group
                float ssum = 0.f, ssq = 0.f;
                float rcur[8], gcur[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    rcur[j] = hasr ? pr[j * rsc] : 0.f;
                    gcur[j] = gate ? __ldg(gate + j) * gs : gs;
                }
#pragma unroll 1
                for (int c0 = 0; c0 < ncols; c0 += 8) {
                    // next chunk's residual / gate loads first: the residual may alias the stores below (in-place update),
                    // so the compiler cannot hoist these loads by itself
                    float rn[8], gn[8];
                    const bool more = c0 + 8 < ncols;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        rn[j] = (hasr && more) ? pr[(c0 + 8 + j) * rsc] : 0.f;
                        gn[j] = (gate && more) ? __ldg(gate + c0 + 8 + j) * gs : gs;
                    }
                    uint32_t r[8];
                    tmem_ld8(tcol + c0, r);
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r[j]), gcur[j], rcur[j] * al);
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) po[(c0 + j) * osc] = v[j];
                    }
                    if (p.stats) {
                        if (gcn >= 8) {
                            // at most one group boundary inside the chunk, at column `left`
                            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float w = ok ? v[j] : 0.f;
                                if (j < left) { a0 += w; a1 = fmaf(w, w, a1); } else { b0 += w; b1 = fmaf(w, w, b1); }
                            }
                            ssum += a0; ssq += a1;
                            left -= 8;
                            if (left <= 0) { flush_group(ssum, ssq, grp); ++grp; ssum = b0; ssq = b1; left += gcn; }
                        } else {
#pragma unroll 1
                            for (int j = 0; j < 8; ++j) {
                                const float w = ok ? v[j] : 0.f;
                                ssum += w; ssq = fmaf(w, w, ssq);
                                if (--left == 0) { flush_group(ssum, ssq, grp); ++grp; ssum = 0.f; ssq = 0.f; left = gcn; }
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { rcur[j] = rn[j]; gcur[j] = gn[j]; }
                }
                if (p.stats && (ssum != 0.f || ssq != 0.f || true)) flush_group(ssum, ssq, min(grp, 7));
            }
            tc_fence_before();
            mbar_arrive(tmem_empty + ab);
            if (++ab == p.acc_bufs) { ab = 0; aphase ^= 1; }
        }
        flush_global();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- operand preparation ---------------------------------------------------------------------------------
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
    v = fminf(fmaxf(v, -60000.f), 60000.f);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

static int tc_ntile_policy = -1;  // AID_TC_NTILE: 0 = 256-wide tiles, 1 = split 256 into 2 x 128 (double-buffered accumulators)
static int tc_ntile(int Cout) {
    if (tc_ntile_policy < 0) tc_ntile_policy = getenv("AID_TC_NTILE") ? atoi(getenv("AID_TC_NTILE")) : 0;
    if (Cout == 256 && tc_ntile_policy == 1) return 128;
    return Cout <= 256 ? Cout : 256;
}

// w[co][ci][kf][kt] (fp32) -> [n-tile][kf][Cin/16][parts: hi|lo][kt][2][Ntile][8] fp16, scaled by 2^10
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, __half* __restrict__ wp, int Ntot, int Ntile, int Cin, int KF, int KT, int parts) {
    const int KS = Cin >> 4;
    const long long total = (long long)Ntot * Cin * KF * KT;  // one (hi, lo) pair per weight
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int e = (int)(r % 8); r /= 8;
        const int n = (int)(r % Ntile); r /= Ntile;
        const int c = (int)(r % 2); r /= 2;
        const int kt = (int)(r % KT); r /= KT;
        const int ks = (int)(r % KS); r /= KS;
        const int kf = (int)(r % KF); r /= KF;
        const int nt = (int)r;
        const int ci = ks * 16 + c * 8 + e, co = nt * Ntile + n;
        const float v = w[(((long long)co * Cin + ci) * KF + kf) * KT + kt] * TC_W_SCALE;
        __half hi, lo;
        split_half(v, hi, lo);
        const long long half_blk = (long long)KT * 2 * Ntile * 8;                       // one of (hi | lo) of a k-step block
        const long long blk = (((long long)nt * KF + kf) * KS + ks) * (parts * half_blk);
        const long long in_blk = (((long long)kt * 2 + c) * Ntile + n) * 8 + e;
        wp[blk + in_blk] = hi;
        if (parts == 2) wp[blk + half_blk + in_blk] = lo;
    }
}

void launch_pack_weight_tc(const float* w, __half* wp, int Cout, int Cin, int KF, int KT, int parts, cudaStream_t s) {
    const long long total = (long long)Cout * Cin * KF * KT;
    pack_weight_tc_kernel<<<(int)min((long long)8192, (total + 255) / 256), 256, 0, s>>>(w, wp, Cout, tc_ntile(Cout), Cin, KF, KT, parts);
    AID_COUNT_LAUNCH(1);
}

__device__ __forceinline__ float gelu_erf_tc(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

// Normalise / modulate / GELU (unet.py:159-163, 479, 482) writing the split-fp16 planar operand:
//   a[b][c/8][f][1+t][c%8] = split(16 * act(x[b,c,f,t] * scale_c)),  pad pixels (index 0 and T+1) = 0
// With stats == nullptr it is a plain layout/precision conversion (scale_c = 1).
// grid: (ceil(rows / rows_per_block), B*C/8), block 256: a block owns a few whole rows of one 8-channel chunk, so the
// per-(clip, channel) scales are computed once per block and every thread streams several pixels (8 coalesced channel-plane
// loads, two 16-byte stores each).
__global__ void __launch_bounds__(256)
gn_act_tc_kernel(TV x, const double* __restrict__ stats, double n_per_group, const float* __restrict__ gamma,
                 const float* __restrict__ affine, long long affine_bstride, int gelu, int PF, int rows_per_block,
                 __half* __restrict__ a_hi, __half* __restrict__ a_lo) {
    const int C8 = x.C >> 3;
    const int c8 = blockIdx.y % C8, b = blockIdx.y / C8;
    const int Tp = x.T + 2, rows_total = x.F + 2 * PF;
    __shared__ float s_scale[8];
    if (threadIdx.x < 8) {
        float sc = 1.f;
        if (stats) {
            const int c = c8 * 8 + threadIdx.x;
            const int g = c / (x.C / 8);
            const double s1 = stats[((long long)b * 8 + g) * 2 + 0], s2 = stats[((long long)b * 8 + g) * 2 + 1];
            double var = (s2 - s1 * s1 / n_per_group) / (n_per_group - 1.0);
            var = var > 0.0 ? var : 0.0;
            const float stdv = (float)sqrt(var);
            const float mod = affine ? (1.f + affine[b * affine_bstride + c]) : 1.f;
            sc = gamma[c] * mod / (stdv + 1e-7f);
        }
        s_scale[threadIdx.x] = sc;
    }
    __syncthreads();
    float sc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = s_scale[j];
    const float* xb = x.p + (long long)b * x.sb + (long long)(c8 * 8) * x.sc;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(rows_total, r0 + rows_per_block);
    for (int fr = r0; fr < r1; ++fr) {
        const int f = fr - PF;
        const bool rowok = f >= 0 && f < x.F;
        const float* src = xb + (long long)(rowok ? f : 0) * x.T;
        const long long obase = (((long long)b * C8 + c8) * rows_total + fr) * Tp;
#pragma unroll 2
        for (int tp = threadIdx.x; tp < Tp; tp += 256) {
            __align__(16) __half hi[8], lo[8];
            const int t = tp - 1;
            if (!rowok || t < 0 || t >= x.T) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { hi[j] = __float2half_rn(0.f); lo[j] = __float2half_rn(0.f); }
            } else {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (long long)j * x.sc + t);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float w = v[j] * sc[j];
                    if (gelu) w = gelu_erf_tc(w);
                    split_half(w * TC_A_SCALE, hi[j], lo[j]);
                }
            }
            *reinterpret_cast<uint4*>(a_hi + (obase + tp) * 8) = *reinterpret_cast<const uint4*>(hi);
            if (a_lo) *reinterpret_cast<uint4*>(a_lo + (obase + tp) * 8) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

// zero rows the tcgen05 kernel needs above and below every [F][T+2] plane (0 in row mode)
int tc_pad_rows(int T, int KF, int dil) {
    if (T % 128 == 0) return 0;
    const int Tp = T + 2;
    return (KF > 1 ? 2 * dil : 0) + (130 + Tp - 1) / Tp + 1;
}

void launch_gn_act_tc(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                      long long affine_bstride, bool gelu, int PF, __half* a_hi, __half* a_lo, cudaStream_t s) {
    const int rows_total = x.F + 2 * PF, Tp = x.T + 2;
    int rpb = max(1, 4096 / Tp);
    // keep at least ~4 blocks per SM in flight
    while (rpb > 1 && (long long)((rows_total + rpb - 1) / rpb) * x.B * (x.C / 8) < device_sm_count() * 4) rpb >>= 1;
    dim3 grid((rows_total + rpb - 1) / rpb, x.B * (x.C / 8));
    gn_act_tc_kernel<<<grid, 256, 0, s>>>(x, stats, (double)n_per_group, gamma, affine, affine_bstride, gelu ? 1 : 0, PF, rpb, a_hi, a_lo);
    AID_COUNT_LAUNCH(1);
}

// plain fp32 NCHW -> split-fp16 planar operand (inputs of proj_in / res_conv / qk)
void launch_to_planar_tc(const TV& x, int PF, __half* a_hi, __half* a_lo, cudaStream_t s) {
    launch_gn_act_tc(x, nullptr, 1, nullptr, nullptr, 0, false, PF, a_hi, a_lo, s);
}

bool conv_tc_supported(int Cin, int Cout, int KF, int KT) {
    const bool k = (KF == 5 && KT == 3) || (KF == 1 && KT == 1);
    return k && Cin % 16 == 0 && Cin >= 16 && Cout % 16 == 0 && Cout >= 16 && (Cout <= 256 || Cout % 256 == 0);
}

// a_hi / a_lo: [B][Cin/8][F + 2*PF][T+2][8] with PF >= tc_pad_rows(T, KF, dil); a_lo == nullptr selects the single-fp16 scheme
// (weights packed with parts = 1)
void launch_conv_tc(const __half* a_hi, const __half* a_lo, int PF, const __half* wp, int B, int Cin, int F, int T, int KF, int KT, int dil,
                    const TV& out, const ConvEpilogue& ep, int num_sms, cudaStream_t s) {
    if (PF < tc_pad_rows(T, KF, dil)) throw CudaError(cudaErrorInvalidValue, "conv_tc: not enough pad rows", __FILE__, __LINE__);
    if (ep.R2.p) throw CudaError(cudaErrorInvalidValue, "conv_tc: R2 is not supported", __FILE__, __LINE__);
    if (!conv_tc_supported(Cin, out.C, KF, KT)) throw CudaError(cudaErrorInvalidValue, "conv_tc: unsupported shape", __FILE__, __LINE__);
    TcConvArgs p{};
    p.a_hi = a_hi; p.a_lo = a_lo; p.w = wp; p.parts = a_lo ? 2 : 1; p.out = out; p.R = ep.R; p.gate = ep.gate; p.gate_bstride = ep.gate_bstride;
    p.alpha = ep.alpha; p.stats = ep.stats;
    p.B = B; p.Cin = Cin; p.Ntot = out.C; p.Ntile = tc_ntile(out.C); p.n_ntiles = out.C / p.Ntile;
    p.F = F; p.T = T; p.Tp = T + 2; p.dil = dil;
    p.KF = KF; p.KT = KT; p.kt_shift = (KT == 1) ? 1 : 0;
    static const int env_kps5 = getenv("AID_TC_KPS5") ? atoi(getenv("AID_TC_KPS5")) : 1;
    static const int env_kps1 = getenv("AID_TC_KPS1") ? atoi(getenv("AID_TC_KPS1")) : 2;
    static const int env_stages = getenv("AID_TC_STAGES") ? atoi(getenv("AID_TC_STAGES")) : 6;
    p.kps = min(min(TC_MAX_KPS, Cin / 16), (KF == 1) ? env_kps1 : env_kps5);
    while (2 * p.kps * p.parts * 2 > 31) --p.kps;   // one producer lane per A plane copy
    p.PF = PF; p.rows_total = F + 2 * PF;
    p.stream = (T % 128 != 0) ? 1 : 0;
    p.tiles_t = (T + 127) / 128;
    p.units_per_b = p.stream ? (F * p.Tp + 127) / 128 : F * p.tiles_t;
    p.n_units = B * p.units_per_b;
    p.n_pairs = (p.n_units + 1) / 2;
    p.n_tiles = p.n_pairs * p.n_ntiles;
    p.b_kstep_bytes = p.parts * KT * 2 * p.Ntile * 16;  // (hi, lo) x kt x 2 chunks x Ntile x 16 B
    p.a_bytes = 2 * p.kps * p.parts * 2 * TC_PLANE;     // 2 units x kps x (hi, lo) x 2 chunks
    p.stage_bytes = p.b_kstep_bytes * p.kps + p.a_bytes;
    p.nstages = min(min(12, env_stages), (224 * 1024) / p.stage_bytes);
    p.ncol_stride = p.Ntile <= 64 ? 64 : (p.Ntile <= 128 ? 128 : 256);
    p.acc_bufs = p.ncol_stride <= 128 ? 2 : 1;
    const size_t smem = (size_t)p.nstages * p.stage_bytes + 256 + 16 * 16 * sizeof(double);  // stages + barriers + per-warp statistics
    static const int dbg = getenv("AID_TC_DEBUG") ? atoi(getenv("AID_TC_DEBUG")) : 0;
    p.dbg = dbg;
    static SmemConfig configured;
    ensure_dyn_smem(conv_tc_kernel, smem, configured);
    const int grid = min(p.n_tiles, num_sms);
    conv_tc_kernel<<<grid, TC_THREADS, smem, s>>>(p);
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
