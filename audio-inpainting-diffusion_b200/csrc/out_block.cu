// Fused "out" block of the decoder / bottleneck (unet.py:452-493 with use_norm, one 1x1 layer, proj_out N -> 2 after it and res_conv
// N -> 2; instantiated at unet.py:690, 719 and called at unet.py:794, 817):
//     x1  = (x + gate * H(GELU(GroupNorm8(x) * gamma * (1 + affine)))) / sqrt 2                 H: 1x1, N -> N
//     blk = (proj_out(x1) + res_conv(x)) / sqrt 2                                              both 1x1, N -> 2
//     out = blk                      (bottleneck)        or        out = (accum + blk) / sqrt 2   (decoder levels)
// Only TWO channels leave the block, and everything after the GELU is linear, so the N x N layer collapses into the projection:
//     proj_out(x1) = (P x + (P diag(gate) H) a) / sqrt 2,     a = GELU(x * s),  s = gamma (1 + affine) / (std + eps)
//     out = W1 x + W2_b a (+ beta accum),      W1 = A (R + P / sqrt 2)  [2 x N],   W2_b = A / sqrt 2 * P diag(gate_b) H  [2 x N per clip]
// with A = 1 / sqrt 2 (no accum) or 1 / 2.  Un-fused this is four launches moving 24 bytes per element of x (operand pass, tcgen05 1x1
// layer with its residual, two thin projections); collapsed it is one pass that reads x once: 4 bytes per element, 2 N^2 multiply-adds
// per clip for W2_b (out_block_prep_kernel), no tensor-core work at all.  It is also closer to the fp32 definition than the un-fused
// conv_mode 2 path: nothing is rounded to fp16 (the un-fused layer rounds a and H).
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace aid {

static constexpr int OB_TH = 256;

// tab[b][c][12] = {w1_0, w1_0, w1_1, w1_1,  w2_0, w2_0, w2_1, w2_1,  cu, cu, ch, ch}: duplicated for the packed fp32x2 pipe (two pixels per
// instruction); w2 carries the 1 / 16 of gelu16_tc2_folded2, which returns 16 GELU(x s) from cu = |s| sqrt(log2 e / 2) and ch = 8 s.
// hw: H K-major [ci][co] (= wp[ci * N + co]); pw, rw: proj_out / res_conv K-major [ci][2].  grid: (ceil(N / 32), B), one warp per 4 channels
// (lanes stride the contraction index: coalesced reads of H, fixed-order shuffle reduction).
__global__ void __launch_bounds__(OB_TH) out_block_prep_kernel(const double* __restrict__ stats, double n_per_group, const float* __restrict__ gamma,
                                                               const float* __restrict__ affine, long long affine_bstride,
                                                               const float* __restrict__ gate, long long gate_bstride, const float* __restrict__ hw,
                                                               const float* __restrict__ pw, const float* __restrict__ rw, int N, float A,
                                                               float* __restrict__ tab) {
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ float s_inv[8];
    extern __shared__ float pg[];      // [N][2]: P[k][n] * gate[n]
    if (threadIdx.x < 8) {             // 1 / (unbiased std + eps) of the 8 groups, the arithmetic of gn_act_tc2_kernel
        const double s1 = stats[((long long)b * 8 + threadIdx.x) * 2 + 0], s2 = stats[((long long)b * 8 + threadIdx.x) * 2 + 1];
        double var = (s2 - s1 * s1 / n_per_group) / (n_per_group - 1.0);
        var = var > 0.0 ? var : 0.0;
        s_inv[threadIdx.x] = 1.f / ((float)sqrt(var) + 1e-7f);
    }
    for (int n = threadIdx.x; n < N; n += OB_TH) {
        const float g = gate ? gate[(long long)b * gate_bstride + n] : 1.f;
        pg[2 * n] = pw[2 * n] * g; pg[2 * n + 1] = pw[2 * n + 1] * g;
    }
    __syncthreads();
    const double kA = (double)A, kAr = (double)A * 0.70710678118654752440;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
        const int c = blockIdx.x * 32 + warp * 4 + j;
        if (c >= N) break;             // (warp-uniform)
        double a0 = 0.0, a1 = 0.0;
        const float* h = hw + (long long)c * N;
        for (int n = lane; n < N; n += 32) { const double hv = (double)__ldg(h + n); a0 += (double)pg[2 * n] * hv; a1 += (double)pg[2 * n + 1] * hv; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
        if (lane == 0) {
            const float mod = affine ? (1.f + affine[(long long)b * affine_bstride + c]) : 1.f;
            const float s = gamma[c] * mod * s_inv[c / (N / 8)];
            const float w10 = (float)(kA * (double)rw[2 * c] + kAr * (double)pw[2 * c]), w11 = (float)(kA * (double)rw[2 * c + 1] + kAr * (double)pw[2 * c + 1]);
            const float w20 = (float)(kAr * a0 / 16.0), w21 = (float)(kAr * a1 / 16.0);
            float4* t = reinterpret_cast<float4*>(tab + ((long long)b * N + c) * 12);
            t[0] = make_float4(w10, w10, w11, w11);
            t[1] = make_float4(w20, w20, w21, w21);
            const float cu = fabsf(s) * 0.84932180028801904272f, ch = 8.f * s;
            t[2] = make_float4(cu, cu, ch, ch);
        }
    }
}

// One thread = 4 consecutive pixels (two fp32x2 pairs), both output channels; streams the N planes of x once.
// grid: (ceil(F * T / 4 / 256), 1, B).  accum may alias out (same element read then written by the same thread).
__global__ void __launch_bounds__(OB_TH) out_block_kernel(TV x, const float* __restrict__ tab, TV out, TV accum, float beta) {
    extern __shared__ __align__(16) float ws[];      // [N][12]
    const int N = x.C, b = blockIdx.z;
    {
        const float4* src = reinterpret_cast<const float4*>(tab + (long long)b * N * 12);
        float4* dst = reinterpret_cast<float4*>(ws);
        for (int i = threadIdx.x; i < N * 3; i += OB_TH) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const long long pq = (long long)blockIdx.x * OB_TH + threadIdx.x;
    if (pq >= ((long long)x.F * x.T) / 4) return;
    const long long pix = pq * 4;                     // rows are contiguous inside a plane
    const float* px = x.p + (long long)b * x.sb + pix;
    const long long sc = x.sc;
    float2 acc[2][2];                                 // [output channel][pixel pair]
#pragma unroll
    for (int k = 0; k < 2; ++k) { acc[k][0] = make_float2(0.f, 0.f); acc[k][1] = make_float2(0.f, 0.f); }
    const float4* w4 = reinterpret_cast<const float4*>(ws);
#pragma unroll 4
    for (int c = 0; c < N; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(px + c * sc));
        const float4 w1 = w4[3 * c], w2 = w4[3 * c + 1], cc = w4[3 * c + 2];
        const float2 v01 = make_float2(v.x, v.y), v23 = make_float2(v.z, v.w);
        const float2 cu = make_float2(cc.x, cc.y), ch = make_float2(cc.z, cc.w);
        const float2 a01 = gelu16_tc2_folded2(v01, cu, ch), a23 = gelu16_tc2_folded2(v23, cu, ch);
        acc[0][0] = __ffma2_rn(make_float2(w1.x, w1.y), v01, acc[0][0]); acc[0][1] = __ffma2_rn(make_float2(w1.x, w1.y), v23, acc[0][1]);
        acc[1][0] = __ffma2_rn(make_float2(w1.z, w1.w), v01, acc[1][0]); acc[1][1] = __ffma2_rn(make_float2(w1.z, w1.w), v23, acc[1][1]);
        acc[0][0] = __ffma2_rn(make_float2(w2.x, w2.y), a01, acc[0][0]); acc[0][1] = __ffma2_rn(make_float2(w2.x, w2.y), a23, acc[0][1]);
        acc[1][0] = __ffma2_rn(make_float2(w2.z, w2.w), a01, acc[1][0]); acc[1][1] = __ffma2_rn(make_float2(w2.z, w2.w), a23, acc[1][1]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float4 r = make_float4(acc[k][0].x, acc[k][0].y, acc[k][1].x, acc[k][1].y);
        if (accum.p) {
            const float4 q = *reinterpret_cast<const float4*>(accum.p + (long long)b * accum.sb + (long long)k * accum.sc + pix);
            r.x = fmaf(beta, q.x, r.x); r.y = fmaf(beta, q.y, r.y); r.z = fmaf(beta, q.z, r.z); r.w = fmaf(beta, q.w, r.w);
        }
        *reinterpret_cast<float4*>(out.p + (long long)b * out.sb + (long long)k * out.sc + pix) = r;
    }
}

static bool ob_aligned16(const TV& v) {
    return v.p == nullptr || ((reinterpret_cast<uintptr_t>(v.p) & 15) == 0 && (v.sb & 3) == 0 && (v.sc & 3) == 0);
}

// x, out and accum must be whole planes (F * T contiguous elements per channel, as every block input / output of the network is)
bool out_block_supported(const TV& x, const TV& out, const TV& accum) {
    return x.C % 8 == 0 && x.C >= 16 && x.C <= 512 && out.C == 2 && ((long long)x.F * x.T) % 4 == 0 && ob_aligned16(x) && ob_aligned16(out) && ob_aligned16(accum) &&
           x.F == out.F && x.T == out.T && (!accum.p || (accum.C == 2 && accum.F == x.F && accum.T == x.T));
}
size_t out_block_scratch_floats(int B, int N) { return (size_t)B * N * 12; }

// hw / pw / rw: fp32 K-major weights of the layer (N x N), proj_out and res_conv (N -> 2); accum.p == nullptr: out = blk, else
// out = (accum + blk) / sqrt 2.  stats: (sum, sumsq) of x per clip and group.
void launch_out_block(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine, long long affine_bstride,
                      const float* gate, long long gate_bstride, const float* hw, const float* pw, const float* rw, const TV& out, const TV& accum,
                      float* scratch, cudaStream_t s) {
    if (!out_block_supported(x, out, accum) || !stats) throw CudaError(cudaErrorInvalidValue, "out_block: unsupported shape", __FILE__, __LINE__);
    const int N = x.C, B = x.B;
    const float A = accum.p ? 0.5f : 0.70710678118654752440f;
    out_block_prep_kernel<<<dim3((N + 31) / 32, B), OB_TH, (size_t)N * 2 * sizeof(float), s>>>(stats, (double)n_per_group, gamma, affine, affine_bstride, gate, gate_bstride, hw, pw, rw,
                                                                         N, A, scratch);
    const long long n4 = ((long long)x.F * x.T) / 4;
    out_block_kernel<<<dim3((unsigned)((n4 + OB_TH - 1) / OB_TH), 1, B), OB_TH, (size_t)N * 12 * sizeof(float), s>>>(x, scratch, out, accum, 0.70710678118654752440f);
    AID_COUNT_LAUNCH(2);
}

}  // namespace aid
