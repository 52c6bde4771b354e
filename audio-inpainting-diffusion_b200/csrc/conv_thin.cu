// HBM-bound "thin" convolutions: very few input or output channels, so there is no GEMM to speak of and the job is
// to stream the big tensor once with wide, coalesced accesses (fp32 CUDA cores, exact).
//   thin-in  (Cin*KF*KT <= 30): init-block proj_in / res_conv 2 -> N (unet.py:412-415, 675), pyr_down_proj 2 -> N 5x3
//                                (unet.py:676, 794), attention proj_out 8 -> N (unet.py:323, 378)
//   thin-out (Cout = 2 or 8)   : out-block proj_out / res_conv N -> 2 (unet.py:412-414, 690, 719), attention proj_in
//                                N -> 8 (unet.py:322, 345)
// Same fused epilogue as the other convolution kernels: out = alpha*(acc*gate + R) + beta*R2, group statistics.
// Weights are the K-major packing wp[(ci*KF*KT + tap)*Cout + co] used by conv_simt.cu.
#include <cstdlib>
#include "common.cuh"

namespace aid {

static constexpr int TH = 256;

__device__ __forceinline__ void thin_stats_flush(float s, float q, int g, double (*sst)[2]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sst[g][0], (double)s); atomicAdd(&sst[g][1], (double)q); }
}

// One thread = PX consecutive pixels of one row, all output channels.  grid: (ceil(F*T/PX / 256), 1, B)
template <int KF, int KT, int CIN, int PX, int CO>
__global__ void __launch_bounds__(TH, CO == 4 ? 3 : 1) conv_thin_in_kernel(TV a, const float* __restrict__ wp, int dil, TV out, ConvEpilogue ep) {
    constexpr int K = CIN * KF * KT;
    // (tried: the multiply-accumulate on the packed fp32x2 pipe with the weights duplicated as (w, w) pairs in shared memory -- one
    // LDS.64 + PX / 2 FFMA2 per tap instead of one LDS + PX FFMA.  Measured slower, 1.89 -> 2.47 ms over the seven pyramid
    // convolutions at B = 8: twice the shared memory per block, fewer resident blocks.)
    extern __shared__ __align__(16) float ws[];  // [K][Cout]
    __shared__ double sst[8][2];
    const int Cout = out.C, F = a.F, T = a.T, b = blockIdx.z;
    for (int i = threadIdx.x; i < K * Cout; i += TH) ws[i] = __ldg(wp + i);
    if (threadIdx.x < 16) sst[threadIdx.x >> 1][threadIdx.x & 1] = 0.0;
    __syncthreads();
    const int tq = T / PX;
    const long long pq = (long long)blockIdx.x * TH + threadIdx.x;
    const bool live = pq < (long long)F * tq;
    const int f = live ? (int)(pq / tq) : 0, t0 = live ? (int)(pq % tq) * PX : 0;

    float in[K][PX];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int kf = 0; kf < KF; ++kf) {
            const int ff = f + (kf - KF / 2) * dil;
            const bool rowok = live && ff >= 0 && ff < F;
            const float* row = a.p + (long long)b * a.sb + (long long)ci * a.sc + (long long)(rowok ? ff : 0) * T;
            float win[PX + KT - 1];
#pragma unroll
            for (int j = 0; j < PX + KT - 1; ++j) {
                const int t = t0 + j - KT / 2;
                win[j] = (rowok && t >= 0 && t < T) ? __ldg(row + t) : 0.f;
            }
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int px = 0; px < PX; ++px) in[(ci * KF + kf) * KT + kt][px] = win[px + kt];
        }

    const int gcn = Cout >= 8 ? Cout / 8 : 1;
    float ssum = 0.f, ssq = 0.f;
    const long long prow = (long long)f * T + t0;
    float* po = out.p + (long long)b * out.sb + prow;
    const float* pr = ep.R.p ? ep.R.p + (long long)b * ep.R.sb + prow : nullptr;
    const float* pr2 = ep.R2.p ? ep.R2.p + (long long)b * ep.R2.sb + prow : nullptr;
    const float* gate = ep.gate ? ep.gate + (long long)b * ep.gate_bstride : nullptr;
    // CO output channels per pass.  Their residual loads are all issued before the first store of the pass: R / R2 may alias out, so the
    // compiler cannot move a load above a store by itself, and with one channel per pass every thread had a single load in flight
    // (16-32 KB per SM, far from what HBM latency x bandwidth needs).  The weights of the CO channels are adjacent: one LDS.128 per tap.
#pragma unroll 1
    for (int co0 = 0; co0 < Cout; co0 += CO) {
        float acc[CO][PX], r1[CO][PX];
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            const int co = co0 + c;
#pragma unroll
            for (int px = 0; px < PX; ++px) { acc[c][px] = 0.f; r1[c][px] = 0.f; }
            if (live && pr) {
                if constexpr (PX == 4) { const float4 q = *reinterpret_cast<const float4*>(pr + (long long)co * ep.R.sc); r1[c][0] = q.x; r1[c][1] = q.y; r1[c][2] = q.z; r1[c][3] = q.w; }
                else if constexpr (PX == 2) { const float2 q = *reinterpret_cast<const float2*>(pr + (long long)co * ep.R.sc); r1[c][0] = q.x; r1[c][1] = q.y; }
                else r1[c][0] = pr[(long long)co * ep.R.sc];
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float w[CO];
            if constexpr (CO == 4) { const float4 q = *reinterpret_cast<const float4*>(ws + k * Cout + co0); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; }
            else w[0] = ws[k * Cout + co0];
#pragma unroll
            for (int c = 0; c < CO; ++c)
#pragma unroll
                for (int px = 0; px < PX; ++px) acc[c][px] = fmaf(w[c], in[k][px], acc[c][px]);
        }
#pragma unroll
        for (int c = 0; c < CO; ++c) {
            const int co = co0 + c;
            if (live) {
                const float g = gate ? gate[co] : 1.f;
                float v[PX], r2[PX];
#pragma unroll
                for (int px = 0; px < PX; ++px) r2[px] = 0.f;
                if (pr2) {      // (the gradient accumulators of the backward pass; not on the forward path)
                    if constexpr (PX == 4) { const float4 q = *reinterpret_cast<const float4*>(pr2 + (long long)co * ep.R2.sc); r2[0] = q.x; r2[1] = q.y; r2[2] = q.z; r2[3] = q.w; }
                    else if constexpr (PX == 2) { const float2 q = *reinterpret_cast<const float2*>(pr2 + (long long)co * ep.R2.sc); r2[0] = q.x; r2[1] = q.y; }
                    else r2[0] = pr2[(long long)co * ep.R2.sc];
                }
#pragma unroll
                for (int px = 0; px < PX; ++px) {
                    const float x = (acc[c][px] * g + r1[c][px]) * ep.alpha + ep.beta * r2[px];
                    v[px] = x;
                    ssum += x; ssq = fmaf(x, x, ssq);
                }
                float* o = po + (long long)co * out.sc;
                if constexpr (PX == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                else if constexpr (PX == 2) *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
                else o[0] = v[0];
            }
            if (ep.stats && ((co + 1) % gcn == 0 || co + 1 == Cout)) {     // (warp-uniform)
                thin_stats_flush(ssum, ssq, min(co / gcn, 7), sst);
                ssum = 0.f; ssq = 0.f;
            }
        }
    }
    if (ep.stats) {
        __syncthreads();
        if (threadIdx.x < 16) {
            const double v = sst[threadIdx.x >> 1][threadIdx.x & 1];
            if (v != 0.0) atomicAdd(ep.stats + (long long)b * 16 + threadIdx.x, v);
        }
    }
}

// Optional input normalisation folded into the weights (attention proj_in, unet.py:343-345: a_in(GroupNorm(x) * gamma * (1 + affine))):
// a bias-free group norm without activation is one scale per (clip, channel), so w'[ci][co] = w[ci][co] * scale[b][ci] and the
// kernel reads the un-normalised tensor -- the separate normalisation pass (4 B read + 4 B written per element) disappears.
struct ThinInScale { const double* stats; double n_per_group; const float* gamma; const float* affine; long long affine_bstride; };

// One thread = 4 consecutive pixels, COUT accumulators each; streams the Cin input planes once.  1x1 only, no statistics.
template <int COUT>
__global__ void __launch_bounds__(TH) conv_thin_out_kernel(TV a, const float* __restrict__ wp, TV out, ConvEpilogue ep, ThinInScale sc) {
    extern __shared__ __align__(16) float ws[];  // [Cin][COUT]
    __shared__ float s_inv[8];
    const int Cin = a.C, F = a.F, T = a.T, b = blockIdx.z;
    if (sc.stats) {
        if (threadIdx.x < 8) {      // 1 / (unbiased std + eps) of the 8 groups, the arithmetic of gn_act_kernel
            const double s1 = sc.stats[((long long)b * 8 + threadIdx.x) * 2 + 0], s2 = sc.stats[((long long)b * 8 + threadIdx.x) * 2 + 1];
            double var = (s2 - s1 * s1 / sc.n_per_group) / (sc.n_per_group - 1.0);
            var = var > 0.0 ? var : 0.0;
            s_inv[threadIdx.x] = 1.f / ((float)sqrt(var) + 1e-7f);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < Cin * COUT; i += TH) {
            const int ci = i / COUT;
            const float mod = sc.affine ? (1.f + sc.affine[b * sc.affine_bstride + ci]) : 1.f;
            ws[i] = __ldg(wp + i) * (sc.gamma[ci] * mod * s_inv[ci / (Cin / 8)]);
        }
    } else {
        for (int i = threadIdx.x; i < Cin * COUT; i += TH) ws[i] = __ldg(wp + i);
    }
    __syncthreads();
    const int tq = T / 4;
    const long long pq = (long long)blockIdx.x * TH + threadIdx.x;
    if (pq >= (long long)F * tq) return;
    const long long prow = (pq / tq) * T + (pq % tq) * 4;
    const float* pa = a.p + (long long)b * a.sb + prow;
    float acc[COUT][4];
#pragma unroll
    for (int co = 0; co < COUT; ++co)
#pragma unroll
        for (int px = 0; px < 4; ++px) acc[co][px] = 0.f;
#pragma unroll 8
    for (int ci = 0; ci < Cin; ++ci) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(pa + (long long)ci * a.sc));
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            const float w = ws[ci * COUT + co];
            acc[co][0] = fmaf(w, x.x, acc[co][0]); acc[co][1] = fmaf(w, x.y, acc[co][1]);
            acc[co][2] = fmaf(w, x.z, acc[co][2]); acc[co][3] = fmaf(w, x.w, acc[co][3]);
        }
    }
    const float* gate = ep.gate ? ep.gate + (long long)b * ep.gate_bstride : nullptr;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        const float g = gate ? gate[co] : 1.f;
        float4 v = make_float4(acc[co][0] * g, acc[co][1] * g, acc[co][2] * g, acc[co][3] * g);
        if (ep.R.p) {
            const float4 r = *reinterpret_cast<const float4*>(ep.R.p + (long long)b * ep.R.sb + (long long)co * ep.R.sc + prow);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        v.x *= ep.alpha; v.y *= ep.alpha; v.z *= ep.alpha; v.w *= ep.alpha;
        if (ep.R2.p) {
            const float4 r = *reinterpret_cast<const float4*>(ep.R2.p + (long long)b * ep.R2.sb + (long long)co * ep.R2.sc + prow);
            v.x += ep.beta * r.x; v.y += ep.beta * r.y; v.z += ep.beta * r.z; v.w += ep.beta * r.w;
        }
        *reinterpret_cast<float4*>(out.p + (long long)b * out.sb + (long long)co * out.sc + prow) = v;
    }
}

// 5x3 (dilation 1) convolution N -> 2: the data gradient of the pyramid projections (unet.py:676, 794), which the general CUDA-core kernel
// served at 0.04 TB/s (15 % of the whole backward).  One thread = 4 consecutive pixels of one row, both output channels; per input
// channel it reads the five tap rows (6 pixels each, served by L1 / L2 after the first row) and does 15 x 4 x 2 FMAs.
__global__ void __launch_bounds__(TH) conv_thin_out53_kernel(TV a, const float* __restrict__ wp, TV out, ConvEpilogue ep) {
    extern __shared__ __align__(16) float ws[];  // [Cin * 15][2]
    const int Cin = a.C, F = a.F, T = a.T, b = blockIdx.z;
    for (int i = threadIdx.x; i < Cin * 30; i += TH) ws[i] = __ldg(wp + i);
    __syncthreads();
    const int tq = T / 4;
    const long long pq = (long long)blockIdx.x * TH + threadIdx.x;
    if (pq >= (long long)F * tq) return;
    const int f = (int)(pq / tq), t0 = (int)(pq % tq) * 4;
    float acc[2][4];
#pragma unroll
    for (int co = 0; co < 2; ++co)
#pragma unroll
        for (int px = 0; px < 4; ++px) acc[co][px] = 0.f;
    const float* pa = a.p + (long long)b * a.sb;
    for (int ci = 0; ci < Cin; ++ci) {
        const float* pc = pa + (long long)ci * a.sc;
        const float* w = ws + ci * 30;
#pragma unroll
        for (int kf = 0; kf < 5; ++kf) {
            const int ff = f + kf - 2;
            if (ff < 0 || ff >= F) continue;
            const float* row = pc + (long long)ff * T + t0;
            const float4 m = __ldg(reinterpret_cast<const float4*>(row));
            const float win[6] = {t0 > 0 ? __ldg(row - 1) : 0.f, m.x, m.y, m.z, m.w, t0 + 4 < T ? __ldg(row + 4) : 0.f};
#pragma unroll
            for (int kt = 0; kt < 3; ++kt) {
                const float w0 = w[(kf * 3 + kt) * 2 + 0], w1 = w[(kf * 3 + kt) * 2 + 1];
#pragma unroll
                for (int px = 0; px < 4; ++px) { acc[0][px] = fmaf(w0, win[px + kt], acc[0][px]); acc[1][px] = fmaf(w1, win[px + kt], acc[1][px]); }
            }
        }
    }
    const long long prow = (long long)f * T + t0;
    const float* gate = ep.gate ? ep.gate + (long long)b * ep.gate_bstride : nullptr;
#pragma unroll
    for (int co = 0; co < 2; ++co) {
        const float g = gate ? gate[co] : 1.f;
        float4 v = make_float4(acc[co][0] * g, acc[co][1] * g, acc[co][2] * g, acc[co][3] * g);
        if (ep.R.p) {
            const float4 r = *reinterpret_cast<const float4*>(ep.R.p + (long long)b * ep.R.sb + (long long)co * ep.R.sc + prow);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        v.x *= ep.alpha; v.y *= ep.alpha; v.z *= ep.alpha; v.w *= ep.alpha;
        if (ep.R2.p) {
            const float4 r = *reinterpret_cast<const float4*>(ep.R2.p + (long long)b * ep.R2.sb + (long long)co * ep.R2.sc + prow);
            v.x += ep.beta * r.x; v.y += ep.beta * r.y; v.z += ep.beta * r.z; v.w += ep.beta * r.w;
        }
        *reinterpret_cast<float4*>(out.p + (long long)b * out.sb + (long long)co * out.sc + prow) = v;
    }
}

static bool aligned16(const TV& v) {
    return v.p == nullptr || ((reinterpret_cast<uintptr_t>(v.p) & 15) == 0 && (v.sb & 3) == 0 && (v.sc & 3) == 0);
}

template <int KF, int KT, int CIN, int PX>
static void launch_thin_in(const TV& a, const float* wp, int dil, const TV& out, const ConvEpilogue& ep, cudaStream_t s) {
    const size_t smem = (size_t)CIN * KF * KT * out.C * sizeof(float);
    static SmemConfig configured;
    const long long n = (long long)a.F * (a.T / PX);
    const dim3 grid((unsigned)((n + TH - 1) / TH), 1, a.B);
    static const bool env_co4 = !(getenv("AID_THIN_CO4") && atoi(getenv("AID_THIN_CO4")) == 0);
    if (out.C % 4 == 0 && env_co4) {
        ensure_dyn_smem(conv_thin_in_kernel<KF, KT, CIN, PX, 4>, smem, configured, 48 * 1024);
        conv_thin_in_kernel<KF, KT, CIN, PX, 4><<<grid, TH, smem, s>>>(a, wp, dil, out, ep);
    } else {
        static SmemConfig configured1;
        ensure_dyn_smem(conv_thin_in_kernel<KF, KT, CIN, PX, 1>, smem, configured1, 48 * 1024);
        conv_thin_in_kernel<KF, KT, CIN, PX, 1><<<grid, TH, smem, s>>>(a, wp, dil, out, ep);
    }
    AID_COUNT_LAUNCH(1);
}

// out[b, 0..8) = W (GroupNorm8(x) * gamma * (1 + affine)) for a 1x1 convolution N -> 8 (attention proj_in) with the normalisation folded
// into per-clip weights; false when the shape is not covered (caller normalises separately)
bool launch_conv_thin_out_normed(const TV& a, const double* stats, long long n_per_group, const float* gamma, const float* affine, long long affine_bstride,
                                 const float* wp, const TV& out, cudaStream_t s) {
    if (!aligned16(a) || !aligned16(out) || out.C != 8 || a.T % 4 != 0 || a.C % 8 != 0 || a.C * out.C * 4 > 48 * 1024) return false;
    const long long n = (long long)a.F * (a.T / 4);
    dim3 grid((unsigned)((n + TH - 1) / TH), 1, a.B);
    conv_thin_out_kernel<8><<<grid, TH, (size_t)a.C * out.C * sizeof(float), s>>>(a, wp, out, ConvEpilogue(), ThinInScale{stats, (double)n_per_group, gamma, affine, affine_bstride});
    AID_COUNT_LAUNCH(1);
    return true;
}

// returns false when the shape is not one of the thin cases (caller falls back to the general CUDA-core kernel)
bool launch_conv_thin(const TV& a, const float* wp, int KF, int KT, int dil, const TV& out, const ConvEpilogue& ep, cudaStream_t s) {
    if (!aligned16(a) || !aligned16(out) || !aligned16(ep.R) || !aligned16(ep.R2)) return false;
    const int T = a.T;
    if (KF == 1 && KT == 1 && !ep.stats && (out.C == 2 || out.C == 8) && T % 4 == 0 && a.C * out.C * 4 <= 48 * 1024) {
        const long long n = (long long)a.F * (T / 4);
        dim3 grid((unsigned)((n + TH - 1) / TH), 1, a.B);
        const size_t smem = (size_t)a.C * out.C * sizeof(float);
        if (out.C == 2) conv_thin_out_kernel<2><<<grid, TH, smem, s>>>(a, wp, out, ep, ThinInScale{});
        else conv_thin_out_kernel<8><<<grid, TH, smem, s>>>(a, wp, out, ep, ThinInScale{});
        AID_COUNT_LAUNCH(1);
        return true;
    }
    if (KF == 5 && KT == 3 && dil == 1 && out.C == 2 && !ep.stats && T % 4 == 0 && a.C * 30 * 4 <= 48 * 1024) {
        const long long n = (long long)a.F * (T / 4);
        conv_thin_out53_kernel<<<dim3((unsigned)((n + TH - 1) / TH), 1, a.B), TH, (size_t)a.C * 30 * sizeof(float), s>>>(a, wp, out, ep);
        AID_COUNT_LAUNCH(1);
        return true;
    }
    if (out.C > 512 || (ep.stats && out.C % 8 != 0)) return false;
    if (KF == 1 && KT == 1 && a.C == 2 && T % 4 == 0) { launch_thin_in<1, 1, 2, 4>(a, wp, 1, out, ep, s); return true; }
    if (KF == 1 && KT == 1 && a.C == 8 && T % 4 == 0) { launch_thin_in<1, 1, 8, 4>(a, wp, 1, out, ep, s); return true; }
    if (KF == 5 && KT == 3 && a.C == 2 && T % 2 == 0) { launch_thin_in<5, 3, 2, 2>(a, wp, dil, out, ep, s); return true; }
    return false;
}

}  // namespace aid
