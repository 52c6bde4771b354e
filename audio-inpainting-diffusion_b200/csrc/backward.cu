// Kernels of the denoiser's vector-Jacobian product with respect to its INPUT (reconstruction guidance, sampler.py:57-113:
// torch.autograd.grad(norm, x) through EDM.denoiser -> unet.py:730-845).  Weights get no gradient on this path.
//   scale_channels      : g * gate[c] * alpha                                  (backward of "* gate", unet.py:482 / 467)
//   gn_bwd_reduce/apply : backward of BiasFreeGroupNorm * (affine+1) [GELU]     (unet.py:147-163, 465, 479, 482)
//   resample_*_adj      : adjoints of the 8-tap down / up resamplers           (unet.py:549-580)
//   bgemm               : small batched strided GEMM (attention backward)      (unet.py:362-371)
//   softmax_rows / _bwd : softmax over keys and its backward
//   cqt_gather_adj, spec_synth_adj : pieces of the CQT analysis / synthesis adjoints (the FFTs and per-band transforms of fft.cu
//                         are reused with swapped window tables)
//   pack_conv_weight_T  : K-major weights of the transposed (data-gradient) convolution: taps flipped, channels swapped
// All fp32 CUDA-core code: exact arithmetic, memory bound.
#include "common.cuh"

namespace aid {

static constexpr int BT = 256;
static constexpr int kChunk = 8192;

__device__ __forceinline__ float gelu_grad(float z) {
    // d/dz [0.5 z (1 + erf(z / sqrt 2))] = 0.5 (1 + erf(z / sqrt 2)) + z exp(-z^2 / 2) / sqrt(2 pi).
    // erf by Abramowitz & Stegun 7.1.26 (1.5e-7 absolute): with u = |z| / sqrt 2 its exponential exp(-u^2) IS exp(-z^2 / 2), so the
    // whole derivative costs one MUFU.EX2 and one MUFU.RCP (erff + expf were ~40 instructions and bound these passes).
    const float az = fabsf(z);
    float t, ex;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, az, 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-0.72134752044448170368f * z * z));      // log2(e) / 2
    float pl = fmaf(t, 1.061405429f, -1.453152027f);
    pl = fmaf(pl, t, 1.421413741f);
    pl = fmaf(pl, t, -0.284496736f);
    pl = fmaf(pl, t, 0.254829592f);
    const float erf_abs = fmaf(-pl * t, ex, 1.f);
    return fmaf(0.5f, copysignf(erf_abs, z), 0.5f) + z * 0.39894228040143267794f * ex;
}

// (sum, sumsq) of a statistics slot -> mean, unbiased std
__device__ __forceinline__ void group_moments(const double* st, double n, float& mean, float& stdv) {
    const double s1 = st[0], s2 = st[1];
    double var = (s2 - s1 * s1 / n) / (n - 1.0);
    var = var > 0.0 ? var : 0.0;
    mean = (float)(s1 / n);
    stdv = (float)sqrt(var);
}

// ---------------------------------------------------------------------------------------------------
// out[b,c,:] = alpha * in[b,c,:] * (vec ? vec[b*vstride + c] : 1)      grid: (chunks, C, B)
__global__ void __launch_bounds__(BT) scale_channels_kernel(TV in, const float* __restrict__ vec, long long vstride, float alpha, TV out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const float sc = alpha * (vec ? vec[(long long)b * vstride + c] : 1.f);
    const long long P = (long long)in.F * in.T;
    const float* src = in.p + (long long)b * in.sb + (long long)c * in.sc;
    float* dst = out.p + (long long)b * out.sb + (long long)c * out.sc;
    const long long start = (long long)blockIdx.x * kChunk, end = min(P, start + kChunk);
    for (long long e = start + threadIdx.x; e < end; e += BT) dst[e] = src[e] * sc;
}
void launch_scale_channels(const TV& in, const float* vec, long long vstride, float alpha, const TV& out, cudaStream_t s) {
    const long long P = (long long)in.F * in.T;
    scale_channels_kernel<<<dim3((unsigned)((P + kChunk - 1) / kChunk), in.C, in.B), BT, 0, s>>>(in, vec, vstride, alpha, out);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// Forward: y = act(x * s_g * w_c), s_g = 1 / (std_g + eps), w_c = gamma_c * (1 + affine_c), act = GELU or identity.
// With gz = g * act'(z):   dx_j = s w gz_j  -  s^2 / ((n-1) std) * D_g * (x_j - mean_g),   D_g = sum_{j in g} gz_j w_c x_j.
// reduce: D[b][grp] += partial sums (double atomics on fp32 block partials).        grid: (chunks, C, B)
__global__ void __launch_bounds__(BT)
gn_bwd_reduce_kernel(TV g, TV x, const double* __restrict__ stats, double npg, const float* __restrict__ gamma, const float* __restrict__ affine,
                     long long abstride, int gelu, double* __restrict__ D, int vec) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int grp = c / (x.C / 8);
    float mean, stdv;
    group_moments(stats + ((long long)b * 8 + grp) * 2, npg, mean, stdv);
    const float s = 1.f / (stdv + 1e-7f);
    const float w = gamma[c] * (affine ? 1.f + affine[(long long)b * abstride + c] : 1.f);
    const float sw = s * w;
    const long long P = (long long)x.F * x.T;
    const float* px = x.p + (long long)b * x.sb + (long long)c * x.sc;
    const float* pg = g.p + (long long)b * g.sb + (long long)c * g.sc;
    const long long start = (long long)blockIdx.x * kChunk, end = min(P, start + kChunk);
    float acc = 0.f;
    if (vec) {      // planes and chunks are multiples of 4 elements, 16-byte aligned
        for (long long e = start + 4 * threadIdx.x; e < end; e += 4 * BT) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(px + e));
            const float4 gv = __ldg(reinterpret_cast<const float4*>(pg + e));
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float gz = gs[i];
                if (gelu) gz *= gelu_grad(xs[i] * sw);
                acc = fmaf(gz * w, xs[i], acc);
            }
        }
    } else
    for (long long e = start + threadIdx.x; e < end; e += BT) {
        const float xv = px[e];
        float gz = pg[e];
        if (gelu) gz *= gelu_grad(xv * sw);
        acc = fmaf(gz * w, xv, acc);
    }
    __shared__ float sh[BT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < BT / 32; ++i) t += (double)sh[i];
        atomicAdd(D + (long long)b * 8 + grp, t);
    }
}
// apply: out = cres * gres + dx   (gres may be null; out may alias gres or g)
__global__ void __launch_bounds__(BT)
gn_bwd_apply_kernel(TV g, TV x, const double* __restrict__ stats, double npg, const float* __restrict__ gamma, const float* __restrict__ affine,
                    long long abstride, int gelu, const double* __restrict__ D, TV gres, float cres, TV out, unsigned int* __restrict__ amax, int vec) {
    const int c = blockIdx.y, b = blockIdx.z;
    float vmax = 0.f;
    const int grp = c / (x.C / 8);
    float mean, stdv;
    group_moments(stats + ((long long)b * 8 + grp) * 2, npg, mean, stdv);
    const float s = 1.f / (stdv + 1e-7f);
    const float w = gamma[c] * (affine ? 1.f + affine[(long long)b * abstride + c] : 1.f);
    const float sw = s * w;
    const float k2 = stdv > 0.f ? (float)((double)s * s / ((npg - 1.0) * stdv) * D[(long long)b * 8 + grp]) : 0.f;
    const long long P = (long long)x.F * x.T;
    const float* px = x.p + (long long)b * x.sb + (long long)c * x.sc;
    const float* pg = g.p + (long long)b * g.sb + (long long)c * g.sc;
    const float* pr = gres.p ? gres.p + (long long)b * gres.sb + (long long)c * gres.sc : nullptr;
    float* po = out.p + (long long)b * out.sb + (long long)c * out.sc;
    const long long start = (long long)blockIdx.x * kChunk, end = min(P, start + kChunk);
    if (vec) {      // (out may alias g or gres: every element is read before it is written, by the same thread)
        for (long long e = start + 4 * threadIdx.x; e < end; e += 4 * BT) {
            const float4 xv = *reinterpret_cast<const float4*>(px + e);
            const float4 gv = *reinterpret_cast<const float4*>(pg + e);
            float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pr) rv = *reinterpret_cast<const float4*>(pr + e);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {gv.x, gv.y, gv.z, gv.w}, rs[4] = {rv.x, rv.y, rv.z, rv.w};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float gz = gs[i];
                if (gelu) gz *= gelu_grad(xs[i] * sw);
                float v = sw * gz - k2 * (xs[i] - mean);
                if (pr) v = fmaf(cres, rs[i], v);
                o[i] = v;
                vmax = fmaxf(vmax, fabsf(v));
            }
            *reinterpret_cast<float4*>(po + e) = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else
    for (long long e = start + threadIdx.x; e < end; e += BT) {
        const float xv = px[e];
        float gz = pg[e];
        if (gelu) gz *= gelu_grad(xv * sw);
        float v = sw * gz - k2 * (xv - mean);
        if (pr) v = fmaf(cres, pr[e], v);
        po[e] = v;
        vmax = fmaxf(vmax, fabsf(v));
    }
    if (amax) {      // max |out| for the per-tensor fp16 scaling of the next data-gradient convolution
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if ((threadIdx.x & 31) == 0 && vmax > 0.f) atomicMax(amax, __float_as_uint(vmax));
    }
}
void launch_gn_bwd(const TV& g, const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                   long long abstride, bool gelu, double* D_scratch, const TV& gres, float cres, const TV& out, cudaStream_t s,
                   unsigned int* amax_out) {
    const long long P = (long long)x.F * x.T;
    const dim3 grid((unsigned)((P + kChunk - 1) / kChunk), x.C, x.B);
    AID_CUDA_CHECK(cudaMemsetAsync(D_scratch, 0, (size_t)x.B * 8 * sizeof(double), s));
    auto al4 = [](const TV& v) { return v.p == nullptr || ((reinterpret_cast<uintptr_t>(v.p) & 15) == 0 && (v.sb & 3) == 0 && (v.sc & 3) == 0); };
    const int vec = (P % 4 == 0 && al4(g) && al4(x) && al4(gres) && al4(out)) ? 1 : 0;
    gn_bwd_reduce_kernel<<<grid, BT, 0, s>>>(g, x, stats, (double)n_per_group, gamma, affine, abstride, gelu ? 1 : 0, D_scratch, vec);
    gn_bwd_apply_kernel<<<grid, BT, 0, s>>>(g, x, stats, (double)n_per_group, gamma, affine, abstride, gelu ? 1 : 0, D_scratch, gres, cres, out, amax_out, vec);
    AID_COUNT_LAUNCH(2);
}

// ---------------------------------------------------------------------------------------------------
__constant__ float c_cubic_b[8] = {-0.01171875f, -0.03515625f, 0.11328125f, 0.43359375f,
                                   0.43359375f,  0.11328125f,  -0.03515625f, -0.01171875f};

// Forward down: y[to] = sum_j k[j] x[reflect(2 to + j - 3)], to in [0, T/2).  Adjoint: gx[n] = sum over the unreflected positions p
// that reflect onto n (p = n, p = -n for 1 <= n <= 3, p = 2(T-1) - n for T-3 <= n <= T-2) of sum_j [p + 3 - j even] k[j] gy[(p+3-j)/2].
// out = beta * out + adj.   grid: (ceil(F*T/256), C, B), T = input (gx) length
__global__ void __launch_bounds__(BT) resample_down_adj_kernel(TV gy, TV gx, float beta) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int T = gx.T, To = gy.T;
    const long long e = (long long)blockIdx.x * BT + threadIdx.x;
    if (e >= (long long)gx.F * T) return;
    const int f = (int)(e / T), n = (int)(e % T);
    const float* row = gy.p + (long long)b * gy.sb + (long long)c * gy.sc + (long long)f * To;
    float acc = 0.f;
    int cand[3] = {n, -n, 2 * (T - 1) - n};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int p = cand[q];
        if (q == 1 && (n == 0 || p < -3)) continue;              // p < 0 must reflect onto n: p = -n, n >= 1; positions start at -3
        if (q == 2 && (p < T || p > T + 2 || p == n)) continue;  // p >= T reflects onto 2(T-1) - p; positions end at T + 2
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int num = p + 3 - j;
            if (num < 0 || (num & 1)) continue;
            const int to = num >> 1;
            if (to < To) acc += c_cubic_b[j] * __ldg(row + to);
        }
    }
    float* o = gx.p + (long long)b * gx.sb + (long long)c * gx.sc + e;
    *o = beta != 0.f ? fmaf(beta, *o, acc) : acc;
}
// Forward up: y[2u] = sum_q k[7-2q] x[reflect(u+q-2)], y[2u+1] = sum_q k[6-2q] x[reflect(u+q-1)], u in [0, T).
// Adjoint over the same preimages p of n (positions run from -2 to T+1).   grid over gx (length T); gy has 2T samples
__global__ void __launch_bounds__(BT) resample_up_adj_kernel(TV gy, TV gx, float beta) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int T = gx.T;
    const long long e = (long long)blockIdx.x * BT + threadIdx.x;
    if (e >= (long long)gx.F * T) return;
    const int f = (int)(e / T), n = (int)(e % T);
    const float* row = gy.p + (long long)b * gy.sb + (long long)c * gy.sc + (long long)f * gy.T;
    float acc = 0.f;
    int cand[3] = {n, -n, 2 * (T - 1) - n};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = cand[k];
        if (k == 1 && (n == 0 || p < -2)) continue;
        if (k == 2 && (p < T || p > T + 1 || p == n)) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ue = p - q + 2, uo = p - q + 1;       // even outputs read position u+q-2, odd ones u+q-1
            if (ue >= 0 && ue < T) acc += c_cubic_b[7 - 2 * q] * __ldg(row + 2 * ue);
            if (uo >= 0 && uo < T) acc += c_cubic_b[6 - 2 * q] * __ldg(row + 2 * uo + 1);
        }
    }
    float* o = gx.p + (long long)b * gx.sb + (long long)c * gx.sc + e;
    *o = beta != 0.f ? fmaf(beta, *o, acc) : acc;
}
void launch_resample_down_adj(const TV& gy, const TV& gx, float beta, cudaStream_t s) {
    resample_down_adj_kernel<<<dim3((unsigned)(((long long)gx.F * gx.T + BT - 1) / BT), gx.C, gx.B), BT, 0, s>>>(gy, gx, beta);
    AID_COUNT_LAUNCH(1);
}
void launch_resample_up_adj(const TV& gy, const TV& gx, float beta, cudaStream_t s) {
    resample_up_adj_kernel<<<dim3((unsigned)(((long long)gx.F * gx.T + BT - 1) / BT), gx.C, gx.B), BT, 0, s>>>(gy, gx, beta);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// C[z](m, n) = alpha * sum_k A[z](m, k) B[z](k, n) + beta * C[z](m, n), every operand with (row, column, batch) element strides.
// 32 x 32 tiles, 256 threads, 4 outputs per thread.  Attention backward only (T <= 256, F <= 512): far from the hot path.
__global__ void __launch_bounds__(256) bgemm_kernel(BGemm p) {
    __shared__ float As[32][33], Bs[32][33];
    const int z = blockIdx.z;
    const float* A = p.A + (long long)z * p.sAz;
    const float* B = p.B + (long long)z * p.sBz;
    float* C = p.C + (long long)z * p.sCz;
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // ty in [0, 8): rows ty, ty+8, ty+16, ty+24
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < p.K; k0 += 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = ty + 8 * r;
            const int m = m0 + i, k = k0 + tx;
            As[i][tx] = (m < p.M && k < p.K) ? A[(long long)m * p.sAm + (long long)k * p.sAk] : 0.f;
            const int kk = k0 + i, n = n0 + tx;
            Bs[i][tx] = (kk < p.K && n < p.N) ? B[(long long)kk * p.sBk + (long long)n * p.sBn] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float bv = Bs[k][tx];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[ty + 8 * r][k], bv, acc[r]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty + 8 * r, n = n0 + tx;
        if (m < p.M && n < p.N) {
            float* o = C + (long long)m * p.sCm + (long long)n * p.sCn;
            *o = p.beta != 0.f ? fmaf(p.beta, *o, p.alpha * acc[r]) : p.alpha * acc[r];
        }
    }
}
void launch_bgemm(const BGemm& p, cudaStream_t s) {
    bgemm_kernel<<<dim3((p.N + 31) / 32, (p.M + 31) / 32, p.batch), 256, 0, s>>>(p);
    AID_COUNT_LAUNCH(1);
}

// rows of S[z][t][:] (length T, contiguous) -> softmax in place.  grid: (rows), one warp per row (block of 32)
__global__ void __launch_bounds__(32) softmax_rows_kernel(float* __restrict__ S, int T) {
    float* row = S + (long long)blockIdx.x * T;
    const int l = threadIdx.x;
    float mx = -INFINITY;
    for (int i = l; i < T; i += 32) mx = fmaxf(mx, row[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = l; i < T; i += 32) { const float e = expf(row[i] - mx); row[i] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int i = l; i < T; i += 32) row[i] *= inv;
}
// gS = scale * P * (gP - sum_j gP_j P_j), in place on gP
__global__ void __launch_bounds__(32) softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ gP, int T, float scale) {
    const float* p = P + (long long)blockIdx.x * T;
    float* g = gP + (long long)blockIdx.x * T;
    const int l = threadIdx.x;
    float dot = 0.f;
    for (int i = l; i < T; i += 32) dot = fmaf(g[i], p[i], dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int i = l; i < T; i += 32) g[i] = scale * p[i] * (g[i] - dot);
}
void launch_softmax_rows(float* S, long long rows, int T, cudaStream_t s) {
    softmax_rows_kernel<<<(unsigned)rows, 32, 0, s>>>(S, T);
    AID_COUNT_LAUNCH(1);
}
void launch_softmax_bwd(const float* P, float* gP, long long rows, int T, float scale, cudaStream_t s) {
    softmax_bwd_kernel<<<(unsigned)rows, 32, 0, s>>>(P, gP, T, scale);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// Adjoint of the CQT analysis' window-and-fold step: Xbar[n] = sum_k wadj_k[m] Y_k[m], n = centre_k + m on the WHOLE circle (the
// analysis reads bins above L/2 for the top bands and wraps), no Hermitian completion.  t.dual must hold the analysis windows
// divided by the band transform length (CqtTables of the adjoint).  accumulate != 0: Xbar += ...      grid: (ceil(L/256), B)
__global__ void __launch_bounds__(256) cqt_gather_adj_kernel(CqtTables t, int oct_lo, int oct_hi, const float2* __restrict__ Y, float2* __restrict__ X, int accumulate) {
    const int n = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
    if (n >= t.L) return;
    const int half = t.L >> 1;
    const float2* y = Y + (long long)b * t.ytotal;
    float2 acc = make_float2(0.f, 0.f);
    int lo, hi;
    if (n <= half) { lo = t.klo[n]; hi = t.khi[n]; }
    else {   // above L/2 only the highest bands reach: scan down while a band's upper edge is beyond n
        hi = t.K - 1; lo = t.K;
        for (int k = t.K - 1; k >= 0; --k) {
            const int lg = t.Lg[k];
            if (t.centre[k] + (lg - lg / 2) <= n) break;
            lo = k;
        }
    }
    for (int k = lo; k <= hi; ++k) {
        const int o = k / t.bins;
        if (o < oct_lo || o > oct_hi) continue;
        const int lg = t.Lg[k], m = n - t.centre[k];
        if (m < -(lg / 2) || m >= lg - lg / 2) continue;
        const int M = t.M[o];
        const float d = __ldg(t.dual + t.woff[k] + (m >= 0 ? m : lg + m));
        const float2 v = y[t.yoff[o] + (long long)(k - o * t.bins) * M + (m >= 0 ? m : M + m)];
        acc.x += v.x * d; acc.y += v.y * d;
    }
    float2* f = X + (long long)b * t.L + n;
    if (accumulate) { acc.x += f->x; acc.y += f->y; }
    *f = acc;
}
void launch_cqt_gather_adj(const CqtTables& t, int B, int oct_lo, int oct_hi, const float2* Y, float2* X, bool accumulate, cudaStream_t s) {
    cqt_gather_adj_kernel<<<dim3((t.L + 255) / 256, B), 256, 0, s>>>(t, oct_lo, oct_hi, Y, X, accumulate ? 1 : 0);
    AID_COUNT_LAUNCH(1);
}

// Adjoint of "Hermitian completion + real part of the inverse FFT" (the tail of the CQT synthesis): G = FFT(gbar) becomes the
// gradient of the half spectrum: x(s / L) at n = 0 and L/2 (real part only), x(2 s / L) for 0 < n < L/2, 0 above L/2.
__global__ void spec_synth_adj_kernel(int L, float2* __restrict__ G, float scale) {
    const int b = blockIdx.y, half = L >> 1;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) {
        float2* g = G + (long long)b * L + n;
        if (n == 0 || n == half) *g = make_float2(g->x * scale, 0.f);
        else if (n < half) *g = make_float2(g->x * 2.f * scale, g->y * 2.f * scale);
        else *g = make_float2(0.f, 0.f);
    }
}
void launch_spec_synth_adj(int B, int L, float2* G, float scale, cudaStream_t s) {
    spec_synth_adj_kernel<<<dim3(device_sm_count() * 4, B), 256, 0, s>>>(L, G, scale);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// Per-tensor scaling of a gradient before its fp16 tensor-core convolution.  amax: max |x| as the bit pattern of a non-negative
// float (atomicMax on unsigned); tc_scale: kappa = 2^floor(log2(192 / amax)) (1 when amax == 0), scal[0] = kappa * pre,
// inv_vec[0..n) = post / kappa.
__global__ void __launch_bounds__(BT) absmax_kernel(TV x, unsigned int* __restrict__ amax) {
    const int c = blockIdx.y, b = blockIdx.z;
    const long long P = (long long)x.F * x.T;
    const float* src = x.p + (long long)b * x.sb + (long long)c * x.sc;
    const long long start = (long long)blockIdx.x * kChunk, end = min(P, start + kChunk);
    float m = 0.f;
    for (long long e = start + threadIdx.x; e < end; e += BT) m = fmaxf(m, fabsf(src[e]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax, __float_as_uint(m));
}
__global__ void tc_scale_kernel(unsigned int* __restrict__ amax, float pre, float post, float* __restrict__ scal, float* __restrict__ inv_vec, int n) {
    const float m = __uint_as_float(*amax);
    float kappa = 1.f;
    if (m > 0.f && isfinite(m)) kappa = exp2f(floorf(log2f(192.f / m)));
    if (threadIdx.x == 0) scal[0] = kappa * pre;
    for (int i = threadIdx.x; i < n; i += blockDim.x) inv_vec[i] = post / kappa;
    __syncthreads();
    if (threadIdx.x == 0) *amax = 0u;     // ready for the next tensor
}
void launch_absmax(const TV& x, unsigned int* amax, cudaStream_t s) {
    const long long P = (long long)x.F * x.T;
    absmax_kernel<<<dim3((unsigned)((P + kChunk - 1) / kChunk), x.C, x.B), BT, 0, s>>>(x, amax);
    AID_COUNT_LAUNCH(1);
}
void launch_tc_scale(unsigned int* amax, float pre, float post, float* scal, float* inv_vec, int n, cudaStream_t s) {
    tc_scale_kernel<<<1, 256, 0, s>>>(amax, pre, post, scal, inv_vec, n);
    AID_COUNT_LAUNCH(1);
}
// w_std[co' = ci][ci' = co][tap'] = wp[(ci*taps + taps-1-tap')*Cout + co]: the transposed, tap-mirrored weight in the standard
// [Cout'][Cin'][KF][KT] layout the tensor-core packers take
__global__ void transpose_weight_std_kernel(const float* __restrict__ wp, float* __restrict__ wstd, int Cout, int Cin, int taps) {
    const long long n = (long long)Cout * Cin * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int tp = (int)(i % taps);
        long long r = i / taps;
        const int co = (int)(r % Cout), ci = (int)(r / Cout);
        wstd[i] = wp[((long long)ci * taps + (taps - 1 - tp)) * Cout + co];
    }
}
void launch_transpose_weight_std(const float* wp, float* wstd, int Cout, int Cin, int taps, cudaStream_t s) {
    const long long n = (long long)Cout * Cin * taps;
    transpose_weight_std_kernel<<<(int)min((long long)4096, (n + 255) / 256), 256, 0, s>>>(wp, wstd, Cout, Cin, taps);
    AID_COUNT_LAUNCH(1);
}

// wp[(ci*taps + tap)*Cout + co] -> wpT[(co*taps + tap')*Cin + ci] with tap' the point-mirrored tap (kf' = KF-1-kf, kt' = KT-1-kt)
__global__ void pack_conv_weight_T_kernel(const float* __restrict__ wp, float* __restrict__ wpT, int Cout, int Cin, int KF, int KT) {
    const int taps = KF * KT;
    const long long n = (long long)Cout * Cin * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        long long r = i / Cin;
        const int tp = (int)(r % taps), co = (int)(r / taps);
        const int tap = taps - 1 - tp;          // (KF-1-kf)*KT + (KT-1-kt) == taps-1 - (kf*KT + kt)
        wpT[i] = wp[((long long)ci * taps + tap) * Cout + co];
    }
}
void launch_pack_conv_weight_T(const float* wp, float* wpT, int Cout, int Cin, int KF, int KT, cudaStream_t s) {
    const long long n = (long long)Cout * Cin * KF * KT;
    pack_conv_weight_T_kernel<<<(int)min((long long)4096, (n + 255) / 256), 256, 0, s>>>(wp, wpT, Cout, Cin, KF, KT);
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
