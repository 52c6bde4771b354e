// Static epilogue of the tcgen05 convolution kernels (conv_tc2.cu, conv_comb.cu): TMEM -> registers -> out = alpha*(acc*gate + R),
// group statistics.  Included after common.cuh and tc_ptx.cuh.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace aid {

static constexpr float T2_A_SCALE = 16.f, T2_W_SCALE = 1024.f, T2_OUT_SCALE = 1.f / (16.f * 1024.f);

// 16 * GELU(x * s) of two elements for the conv_mode 2 operand (exact-erf definition to ~5e-7 absolute, Abramowitz & Stegun 7.1.26,
// MUFU rcp / ex2), with the per-channel scale folded into cu = |s| * sqrt(log2(e) / 2) and ch = 8 s.
// Two elements at once with the packed fp32x2 instructions of sm_100 (FMUL2 / FFMA2): the same IEEE operations per lane (|x| cu
// == |x cu| for cu >= 0, the polynomial is evaluated with negated coefficients so that 1 - pl ex is one FFMA2), 9 instead of 14
// instructions per element -- this pass is bound by instruction issue before it is bound by HBM.
__device__ __forceinline__ float2 gelu16_tc2_folded2(float2 x, float2 cu, float2 ch) {
    const float2 xc = __fmul2_rn(x, cu);
    float2 t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(0.3275911f * 0.83255461115769775635f, fabsf(xc.x), 1.f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(0.3275911f * 0.83255461115769775635f, fabsf(xc.y), 1.f)));
    float2 pl = __ffma2_rn(t, make_float2(-1.061405429f, -1.061405429f), make_float2(1.453152027f, 1.453152027f));
    pl = __ffma2_rn(pl, t, make_float2(-1.421413741f, -1.421413741f));
    pl = __ffma2_rn(pl, t, make_float2(0.284496736f, 0.284496736f));
    pl = __ffma2_rn(pl, t, make_float2(-0.254829592f, -0.254829592f));
    pl = __fmul2_rn(pl, t);                                  // = -(polynomial)
    const float2 sq = __fmul2_rn(xc, xc);
    float2 ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(-sq.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(-sq.y));
    const float2 erf_abs = __ffma2_rn(pl, ex, make_float2(1.f, 1.f));
    const float2 h = __fmul2_rn(x, ch);
    return make_float2(fmaf(fabsf(h.x), erf_abs.x, h.x), fmaf(fabsf(h.y), erf_abs.y, h.y));
}
// two operand values (already x16) -> packed fp16x2, saturating to the finite range
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}


// ---- fast epilogue ------------------------------------------------------------------------------------------------------------
// The generic epilogue below handles every layout and group width with run-time bookkeeping; its inner loop compiled to ~20
// instructions per output element with indirect branches in the statistics (ncu: 148 M warp instructions for 134 M outputs of a
// 64-channel layer, issue slots 38 % busy with two epilogue warps per scheduler) and bounded the layers with <= 128 couts.
// This version covers the shapes of the paper networks (NCHW out / R, 8 epilogue warps, a warp's columns = NB batches of BW
// columns whose boundaries coincide with the statistics groups of GCN columns) with everything static: the batch loop is
// unrolled, residuals ping-pong between two register arrays (no copies), group sums are fixed trees, stores and loads use
// running pointers.  Same arithmetic contract as the generic one: per-unit fp32 partial sums -> private double accumulators in
// shared memory -> one double atomic per (clip, group, warp).
// The unit sequence comes from an iterator `IT`: bool next(EpiUnit&) fills the next unit of this CTA (false: none left).
struct EpiArgs {
    TV out, R;                      // NCHW views; R.p == nullptr: no residual
    const float* gate; long long gate_bstride;
    float alpha; double* stats;
    int Ntile, n_ntiles;
    // RES = 1 (conv_init.cu): the residual is not loaded but synthesised from a 2-channel tensor x2 of the same [F][T] geometry,
    // R[c] = x2[0] * c0tab[b][c] + x2[1] * c1tab[b][c]  (tables per clip, tab_bstride apart)
    TV x2 = TV(); const float* c0tab = nullptr; const float* c1tab = nullptr; long long tab_bstride = 0;
};
struct EpiUnit {
    int b, nt;                      // clip, n-tile
    long long pix;                  // this thread's pixel inside the clip's [F][T] plane (any valid pixel when !ok: read, never written)
    bool ok;                        // the thread owns a real pixel of an existing unit
    uint32_t tcol;                  // accumulator column offset of the unit inside the TMEM allocation
    int ab; uint32_t aph;           // accumulator handshake: barrier index and phase
    bool first, last;               // first / last unit that uses this accumulator handshake (wait tmem_full / arrive tmem_empty)
};
template <bool CG2, int BW, int NB, int GCN, class IT, int EWARPS = 8, int RES = 0, int GSTRIDE = 128>
__device__ __forceinline__ void epilogue_fast(const EpiArgs& p, IT& it, const int e, const int lane, const uint32_t tmem_base, float* gsm_base,
                                              double* sacc_base, uint64_t* tmem_full, uint64_t* tmem_empty) {
    constexpr int NCOLS = BW * NB;                  // columns of this warp (the n-tile split over EWARPS / 4 column groups)
    constexpr int SST = EWARPS * 32;                // stride of the double accumulators: [k][epilogue threads]
    constexpr bool DIRECT = (NCOLS / GCN) > 4;      // many groups per warp: batch sums go straight to the shared-memory doubles
    constexpr int NG = NCOLS / GCN;                 // statistics groups they span
    constexpr int GPB = BW / GCN;                   // groups per batch
    static_assert(NB % 2 == 0 && BW % 4 == 0 && BW <= 32 && BW % GCN == 0 && NG >= 1 && NG <= 8 && 2 * NG * SST * 8 <= 16384, "epilogue_fast shape");
    const int q = e & 3, cw = e >> 2, cbeg = cw * NCOLS;
    const float al = p.alpha, gs = T2_OUT_SCALE * p.alpha;
    const uint32_t osc = (uint32_t)p.out.sc, rsc = (uint32_t)p.R.sc;   // plane strides; a clip's tensor has < 2^31 elements
    const bool has_r = p.R.p != nullptr, has_gate = p.gate != nullptr, do_stats = p.stats != nullptr;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbeg;
    static_assert(RES == 0 || GSTRIDE >= 3 * BW * NB, "epilogue_fast: table stride");
    float* gsm = gsm_base + e * GSTRIDE;      // [gate NCOLS] (RES = 1: [c0 NCOLS][c1 NCOLS] behind it)
    const uint32_t gsm_addr = smem_u32(gsm);
    double* sacc = sacc_base + (e * 32 + lane);   // [k][SST threads]
    constexpr int NSQ = DIRECT ? 1 : NG;
    float S[NSQ], Q[NSQ];
#pragma unroll
    for (int k = 0; k < NSQ; ++k) { S[k] = 0.f; Q[k] = 0.f; }
    if (do_stats) {
#pragma unroll
        for (int k = 0; k < 2 * NG; ++k) sacc[k * SST] = 0.0;
    }
    int b_cur = -1, nt_cur = 0, gate_key = -2;
    auto flush_stats = [&]() {
        if (do_stats && b_cur >= 0) {
            double v[2 * NG];
#pragma unroll
            for (int k = 0; k < 2 * NG; ++k) {
                v[k] = sacc[k * SST]; sacc[k * SST] = 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            }
            if (lane == 0) {
                const int g0 = (nt_cur * p.Ntile + cbeg) / GCN;
#pragma unroll
                for (int k = 0; k < 2 * NG; ++k) atomicAdd(p.stats + (long long)b_cur * 16 + (g0 + (k >> 1)) * 2 + (k & 1), v[k]);
            }
        }
    };
    // descriptor of the next unit (its first batch of residuals is prefetched during our last batch)
    float* n_po = nullptr; const float* n_pr = nullptr; uint32_t n_oo = 0, n_ro = 0; uint32_t n_tcol = 0, n_aph = 0; int n_b = 0, n_nt = 0, n_ab = 0;
    bool n_valid = false, n_ok = false, n_first = false, n_last = false;
    float n_x0 = 0.f, n_x1 = 0.f;      // RES = 1: the unit's two input channels at this thread's pixel
    auto advance = [&]() {
        EpiUnit u;
        n_valid = it.next(u);
        if (!n_valid) return;
        n_ok = u.ok; n_nt = u.nt; n_b = u.b; n_ab = u.ab; n_aph = u.aph; n_first = u.first; n_last = u.last;
        // warp-uniform clip base pointers + 32-bit per-thread element offsets: one IMAD.WIDE per access
        const uint32_t co0 = (uint32_t)(n_nt * p.Ntile + cbeg);
        n_po = p.out.p + (long long)u.b * p.out.sb; n_oo = co0 * osc + (uint32_t)u.pix;
        n_pr = p.R.p + (long long)u.b * p.R.sb; n_ro = co0 * rsc + (uint32_t)u.pix;
        n_tcol = tq + u.tcol;
        if constexpr (RES == 1) {
            const float* q = p.x2.p + (long long)u.b * p.x2.sb + u.pix;
            n_x0 = __ldg(q); n_x1 = __ldg(q + p.x2.sc);
        }
    };
    float ra[BW], rb[BW];
    auto load_batch = [&](float (&dst)[BW], const float* src, uint32_t off) {
        if (has_r) {
#pragma unroll
            for (int j = 0; j < BW; ++j) dst[j] = src[off + (uint32_t)j * rsc];     // may alias out: plain loads
        } else {
#pragma unroll
            for (int j = 0; j < BW; ++j) dst[j] = 0.f;
        }
    };
    advance();
    if (RES == 0 && n_valid) load_batch(ra, n_pr, n_ro);
    while (n_valid) {
        float* c_po = n_po; const float* c_pr = n_pr; const uint32_t c_oo = n_oo, c_ro = n_ro;
        const uint32_t c_tcol = n_tcol, c_aph = n_aph; const int c_b = n_b, c_nt = n_nt, c_ab = n_ab;
        const bool c_ok = n_ok, c_first = n_first, c_last = n_last;
        const float c_x0 = n_x0, c_x1 = n_x1;
        advance();
        if (c_first) { mbar_wait(tmem_full + c_ab, c_aph); tc_fence_after(); }
        if (c_b != b_cur || c_nt != nt_cur) { flush_stats(); b_cur = c_b; nt_cur = c_nt; }
        const int gkey = p.gate_bstride ? c_b * p.n_ntiles + c_nt : c_nt;
        if (gkey != gate_key) {
            gate_key = gkey;
            __syncwarp();
            for (int k = lane; k < NCOLS; k += 32)
                gsm[k] = has_gate ? __ldg(p.gate + (long long)c_b * p.gate_bstride + c_nt * p.Ntile + cbeg + k) * gs : gs;
            if constexpr (RES == 1) {
                for (int k = lane; k < NCOLS; k += 32) {
                    gsm[NCOLS + k] = __ldg(p.c0tab + (long long)c_b * p.tab_bstride + cbeg + k);
                    gsm[2 * NCOLS + k] = __ldg(p.c1tab + (long long)c_b * p.tab_bstride + cbeg + k);
                }
            }
            __syncwarp();
        }
        const float m = c_ok ? 1.f : 0.f;
#pragma unroll
        for (int bi = 0; bi < NB; ++bi) {
            float (&cur)[BW] = (bi & 1) ? rb : ra;
            float (&nxt)[BW] = (bi & 1) ? ra : rb;
            if constexpr (RES == 1) {       // synthesise this batch's residual from the unit's two input values and the per-clip tables
#pragma unroll
                for (int j = 0; j < BW; j += 4) {
                    float a0[4], a1[4];
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a0[0]), "=f"(a0[1]), "=f"(a0[2]), "=f"(a0[3]) : "r"(gsm_addr + (uint32_t)(NCOLS + bi * BW + j) * 4u));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a1[0]), "=f"(a1[1]), "=f"(a1[2]), "=f"(a1[3]) : "r"(gsm_addr + (uint32_t)(2 * NCOLS + bi * BW + j) * 4u));
#pragma unroll
                    for (int i = 0; i < 4; ++i) cur[j + i] = fmaf(c_x0, a0[i], c_x1 * a1[i]);
                }
                (void)nxt;
            } else {
                if (bi + 1 < NB) load_batch(nxt, c_pr, c_ro + (uint32_t)((bi + 1) * BW) * rsc);
                else if (n_valid) load_batch(nxt, n_pr, n_ro);
            }
            uint32_t acc[BW];
            if constexpr (BW == 32) tmem_ld32_nowait(c_tcol + bi * BW, acc);
            else if constexpr (BW == 24) { tmem_ld16_nowait(c_tcol + bi * BW, acc); tmem_ld8p_nowait(c_tcol + bi * BW + 16, acc + 16); }
            else if constexpr (BW == 16) tmem_ld16_nowait(c_tcol + bi * BW, acc);
            else if constexpr (BW == 12) { tmem_ld8p_nowait(c_tcol + bi * BW, acc); tmem_ld4p_nowait(c_tcol + bi * BW + 8, acc + 8); }
            else tmem_ld8p_nowait(c_tcol + bi * BW, acc);
            float g[BW];
#pragma unroll
            for (int j = 0; j < BW; j += 4)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g[j]), "=f"(g[j + 1]), "=f"(g[j + 2]), "=f"(g[j + 3]) : "r"(gsm_addr + (uint32_t)(bi * BW + j) * 4u));
            tmem_wait_ld();
            // packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 of sm_100): half the math instructions, the same IEEE results per lane
            float2 v2[BW / 2];
#pragma unroll
            for (int j = 0; j < BW / 2; ++j)
                v2[j] = __ffma2_rn(make_float2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1])), make_float2(g[2 * j], g[2 * j + 1]),
                                   __fmul2_rn(make_float2(cur[2 * j], cur[2 * j + 1]), make_float2(al, al)));
            if (c_ok) {
                const uint32_t ob = c_oo + (uint32_t)(bi * BW) * osc;
#pragma unroll
                for (int j = 0; j < BW / 2; ++j) { c_po[ob + (uint32_t)(2 * j) * osc] = v2[j].x; c_po[ob + (uint32_t)(2 * j + 1) * osc] = v2[j].y; }     // 32-bit offsets: one IMAD.WIDE per store.  (Streaming stores, st.global.cs, measured 2-5 % slower on every layer.)
            }
            if (do_stats) {
#pragma unroll
                for (int gg = 0; gg < GPB; ++gg) {
                    // fixed order over the GCN columns of the group: two interleaved packed chains, then the four lanes
                    constexpr int H = GCN / 2;       // float2 values of the group (even)
                    const float2* w = v2 + gg * H;
                    float2 sa = w[0], sb = w[1], qa = __fmul2_rn(w[0], w[0]), qb = __fmul2_rn(w[1], w[1]);
#pragma unroll
                    for (int k = 2; k < H; k += 2) {
                        sa = __fadd2_rn(sa, w[k]); sb = __fadd2_rn(sb, w[k + 1]);
                        qa = __ffma2_rn(w[k], w[k], qa); qb = __ffma2_rn(w[k + 1], w[k + 1], qb);
                    }
                    const float2 s2 = __fadd2_rn(sa, sb), q2 = __fadd2_rn(qa, qb);
                    if constexpr (DIRECT) {
                        sacc[(2 * (bi * GPB + gg)) * SST] += (double)((s2.x + s2.y) * m);
                        sacc[(2 * (bi * GPB + gg) + 1) * SST] += (double)((q2.x + q2.y) * m);
                    } else {
                        S[bi * GPB + gg] = fmaf(s2.x + s2.y, m, S[bi * GPB + gg]);
                        Q[bi * GPB + gg] = fmaf(q2.x + q2.y, m, Q[bi * GPB + gg]);
                    }
                }
            }
        }
        if constexpr (!DIRECT) {
            if (do_stats) {
#pragma unroll
                for (int k = 0; k < NG; ++k) { sacc[(2 * k) * SST] += (double)S[k]; sacc[(2 * k + 1) * SST] += (double)Q[k]; S[k] = 0.f; Q[k] = 0.f; }
            }
        }
        if (c_last) {      // one arrival per warp: 256 per-thread arrivals on one mbarrier serialise in the shared-memory pipe
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CG2) mbar_arrive_cluster(tmem_empty + c_ab, 0u); else mbar_arrive(tmem_empty + c_ab); }
        }
    }
    flush_stats();
}

}  // namespace aid
