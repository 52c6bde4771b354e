// HBM-bound element-wise / reduction kernels of the denoiser path (sm_100a).
//   group_stats   : per-(clip, group) sum and sum of squares            unet.py:147-155 (std over C/8*F*T)
//   gn_act        : x / (std+eps) * gamma * (affine+1) [, GELU]          unet.py:159-163, 465, 479, 482
//   combine       : alpha*a + beta*b with optional statistics           unet.py:491 (identity res_conv), copies
//   resample      : 8-tap 'cubic' FIR down/up by 2 along T, reflect pad  unet.py:549-580
//   embedding     : RFF + 3-layer MLP of c_noise                         unet.py:184-211
//   mod_vectors   : every adaLN affine/gate Linear of the network       unet.py:430-431, 442-443, 36-40
//   edm_*         : sampler element-wise updates                         sampler.py:214, 141-147, 230-251
#include "common.cuh"

namespace aid {

static constexpr int kThreads = 256;
static constexpr int kElemsPerCta = 8192;  // 256 threads x 8 float4

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two doubles; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
    __shared__ double sh[2][kThreads / 32];
    a = warp_sum(a); b = warp_sum(b);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = a; sh[1][w] = b; }
    __syncthreads();
    if (w == 0) {
        a = l < (int)(blockDim.x >> 5) ? sh[0][l] : 0.0;
        b = l < (int)(blockDim.x >> 5) ? sh[1][l] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
    }
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------------
// grid: (chunks_per_plane * gc, 8 groups, B)
__global__ void __launch_bounds__(kThreads) group_stats_kernel(TV x, int gc, int chunks_per_plane, double* stats) {
    const int g = blockIdx.y, b = blockIdx.z;
    const int cl = blockIdx.x / chunks_per_plane, chunk = blockIdx.x % chunks_per_plane;
    const long long P = (long long)x.F * x.T;
    const float* base = x.p + (long long)b * x.sb + (long long)(g * gc + cl) * x.sc;
    const long long start = (long long)chunk * kElemsPerCta;
    const long long end = min(P, start + kElemsPerCta);
    float s = 0.f, q = 0.f;
    if ((P & 3) == 0 && aligned16(base)) {
        const float4* b4 = reinterpret_cast<const float4*>(base);
        for (long long e = (start >> 2) + threadIdx.x; e < (end >> 2); e += kThreads) {
            float4 v = __ldg(b4 + e);
            s += (v.x + v.y) + (v.z + v.w);
            q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    } else {
        for (long long e = start + threadIdx.x; e < end; e += kThreads) {
            float v = __ldg(base + e);
            s += v; q += v * v;
        }
    }
    double ds = s, dq = q;
    block_sum2(ds, dq);
    if (threadIdx.x == 0) {
        atomicAdd(stats + ((long long)b * 8 + g) * 2 + 0, ds);
        atomicAdd(stats + ((long long)b * 8 + g) * 2 + 1, dq);
    }
}

void launch_group_stats(const TV& x, double* stats, cudaStream_t s) {
    const int gc = x.C / 8;
    const long long P = (long long)x.F * x.T;
    const int cpp = (int)((P + kElemsPerCta - 1) / kElemsPerCta);
    dim3 grid(cpp * gc, 8, x.B);
    group_stats_kernel<<<grid, kThreads, 0, s>>>(x, gc, cpp, stats);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

// grid: (chunks_per_plane, C, B)
__global__ void __launch_bounds__(kThreads)
gn_act_kernel(TV x, const double* __restrict__ stats, double n_per_group, const float* __restrict__ gamma,
              const float* __restrict__ affine, long long affine_bstride, int gelu, TV out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int g = c / (x.C / 8);
    __shared__ float s_scale;
    if (threadIdx.x == 0) {
        const double s1 = stats[((long long)b * 8 + g) * 2 + 0], s2 = stats[((long long)b * 8 + g) * 2 + 1];
        double var = (s2 - s1 * s1 / n_per_group) / (n_per_group - 1.0);  // unbiased, torch.std default
        var = var > 0.0 ? var : 0.0;
        const float stdv = (float)sqrt(var);
        const float mod = affine ? (1.f + affine[b * affine_bstride + c]) : 1.f;
        s_scale = gamma[c] * mod / (stdv + 1e-7f);
    }
    __syncthreads();
    const float sc = s_scale;
    const long long P = (long long)x.F * x.T;
    const float* src = x.p + (long long)b * x.sb + (long long)c * x.sc;
    float* dst = out.p + (long long)b * out.sb + (long long)c * out.sc;
    const long long start = (long long)blockIdx.x * kElemsPerCta;
    const long long end = min(P, start + kElemsPerCta);
    if ((P & 3) == 0 && aligned16(src) && aligned16(dst)) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (long long e = (start >> 2) + threadIdx.x; e < (end >> 2); e += kThreads) {
            float4 v = __ldg(s4 + e);
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
            if (gelu) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
            d4[e] = v;
        }
    } else {
        for (long long e = start + threadIdx.x; e < end; e += kThreads) {
            float v = __ldg(src + e) * sc;
            dst[e] = gelu ? gelu_erf(v) : v;
        }
    }
}

void launch_gn_act(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                   long long affine_bstride, bool gelu, const TV& out, cudaStream_t s) {
    const long long P = (long long)x.F * x.T;
    dim3 grid((unsigned)((P + kElemsPerCta - 1) / kElemsPerCta), x.C, x.B);
    gn_act_kernel<<<grid, kThreads, 0, s>>>(x, stats, (double)n_per_group, gamma, affine, affine_bstride, gelu ? 1 : 0, out);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// out = alpha*a + beta*b   (b.p may be null).  grid: (chunks_per_plane, C, B)
__global__ void __launch_bounds__(kThreads) combine_kernel(TV a, TV bb, float alpha, float beta, TV out, double* stats) {
    const int c = blockIdx.y, b = blockIdx.z;
    const long long P = (long long)a.F * a.T;
    const float* pa = a.p + (long long)b * a.sb + (long long)c * a.sc;
    const float* pb = bb.p ? bb.p + (long long)b * bb.sb + (long long)c * bb.sc : nullptr;
    float* po = out.p + (long long)b * out.sb + (long long)c * out.sc;
    const long long start = (long long)blockIdx.x * kElemsPerCta;
    const long long end = min(P, start + kElemsPerCta);
    float s = 0.f, q = 0.f;
    if ((P & 3) == 0 && aligned16(pa) && aligned16(po) && (!pb || aligned16(pb))) {
        for (long long e = (start >> 2) + threadIdx.x; e < (end >> 2); e += kThreads) {
            float4 v = __ldg(reinterpret_cast<const float4*>(pa) + e);
            v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
            if (pb) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(pb) + e);
                v.x += beta * w.x; v.y += beta * w.y; v.z += beta * w.z; v.w += beta * w.w;
            }
            reinterpret_cast<float4*>(po)[e] = v;
            s += (v.x + v.y) + (v.z + v.w);
            q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    } else {
        for (long long e = start + threadIdx.x; e < end; e += kThreads) {
            float v = alpha * pa[e];
            if (pb) v += beta * pb[e];
            po[e] = v;
            s += v; q += v * v;
        }
    }
    if (stats) {
        double ds = s, dq = q;
        block_sum2(ds, dq);
        if (threadIdx.x == 0) {
            const int g = c / (a.C / 8);
            atomicAdd(stats + ((long long)b * 8 + g) * 2 + 0, ds);
            atomicAdd(stats + ((long long)b * 8 + g) * 2 + 1, dq);
        }
    }
}

void launch_combine(const TV& a, const TV& b, float alpha, float beta, const TV& out, double* stats, cudaStream_t s) {
    const long long P = (long long)a.F * a.T;
    dim3 grid((unsigned)((P + kElemsPerCta - 1) / kElemsPerCta), a.C, a.B);
    combine_kernel<<<grid, kThreads, 0, s>>>(a, b, alpha, beta, out, stats);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
__constant__ float c_cubic[8] = {-0.01171875f, -0.03515625f, 0.11328125f, 0.43359375f,
                                 0.43359375f,  0.11328125f,  -0.03515625f, -0.01171875f};

__device__ __forceinline__ int reflect(int n, int T) {
    n = n < 0 ? -n : n;
    return n >= T ? 2 * (T - 1) - n : n;
}

// y[to] = sum_j k[j] * x[reflect(2*to + j - 3)]          grid: (ceil(F*To/256), C, B)   (scalar fallback)
__global__ void __launch_bounds__(kThreads) resample_down_kernel(TV x, TV out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int To = out.T;
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= (long long)out.F * To) return;
    const int f = (int)(e / To), to = (int)(e % To);
    const float* row = x.p + (long long)b * x.sb + (long long)c * x.sc + (long long)f * x.T;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += c_cubic[j] * __ldg(row + reflect(2 * to + j - 3, x.T));
    out.p[(long long)b * out.sb + (long long)c * out.sc + e] = acc;
}

// 4 outputs per thread: 16 aligned inputs (four float4 loads) in the interior, reflected scalar loads at the row ends.
// grid: (ceil(F*To/4 / 256), C, B); requires To % 4 == 0 and 16-byte aligned rows.
__global__ void __launch_bounds__(kThreads) resample_down4_kernel(TV x, TV out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int To = out.T, T = x.T, tq = To >> 2;
    const int q = blockIdx.x * kThreads + threadIdx.x;
    if (q >= out.F * tq) return;
    const int f = q / tq, to = (q - f * tq) << 2;
    const float* row = x.p + (long long)b * x.sb + (long long)c * x.sc + (long long)f * T;
    float w[16];  // x[2*to - 4 .. 2*to + 11]
    if (to > 0 && to + 4 < To) {
        const float4* r4 = reinterpret_cast<const float4*>(row + 2 * to - 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float4 v = __ldg(r4 + i); w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = __ldg(row + reflect(2 * to - 4 + i, T));
    }
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += c_cubic[j] * w[2 * i + j + 1];  // x[2*(to+i) + j - 3] = w[2i + j + 1]
        y[i] = acc;
    }
    *reinterpret_cast<float4*>(out.p + (long long)b * out.sb + (long long)c * out.sc + (long long)f * To + to) = make_float4(y[0], y[1], y[2], y[3]);
}

// conv_transpose1d(reflect_pad(x, 2), k, stride 2, padding 7):
//   y[2u]   = sum_q k[7-2q] * x[reflect(u+q-2)],   y[2u+1] = sum_q k[6-2q] * x[reflect(u+q-1)],  q = 0..3
__global__ void __launch_bounds__(kThreads) resample_up_kernel(TV x, TV out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int To = out.T;
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= (long long)out.F * To) return;
    const int f = (int)(e / To), m = (int)(e % To);
    const float* row = x.p + (long long)b * x.sb + (long long)c * x.sc + (long long)f * x.T;
    const int u = m >> 1, odd = m & 1;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float w = odd ? c_cubic[6 - 2 * q] : c_cubic[7 - 2 * q];
        acc += w * __ldg(row + reflect(u + q - 2 + odd, x.T));
    }
    out.p[(long long)b * out.sb + (long long)c * out.sc + e] = acc;
}

// 4 outputs (m = 4v .. 4v+3) per thread from the 6 inputs x[2v-2 .. 2v+3]; one float4 store.  Requires To % 4 == 0.
__global__ void __launch_bounds__(kThreads) resample_up4_kernel(TV x, TV out) {
    const int c = blockIdx.y, b = blockIdx.z;
    const int To = out.T, T = x.T, tq = To >> 2;
    const int qd = blockIdx.x * kThreads + threadIdx.x;
    if (qd >= out.F * tq) return;
    const int f = qd / tq, v = qd - f * tq;
    const float* row = x.p + (long long)b * x.sb + (long long)c * x.sc + (long long)f * T;
    float w[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = __ldg(row + reflect(2 * v - 2 + i, T));
    float y[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        y[0] += c_cubic[7 - 2 * q] * w[q];       // u = 2v   (even):  x[u+q-2] = w[q]
        y[1] += c_cubic[6 - 2 * q] * w[q + 1];   // u = 2v   (odd) :  x[u+q-1] = w[q+1]
        y[2] += c_cubic[7 - 2 * q] * w[q + 1];   // u = 2v+1 (even):  x[u+q-2] = w[q+1]
        y[3] += c_cubic[6 - 2 * q] * w[q + 2];   // u = 2v+1 (odd) :  x[u+q-1] = w[q+2]
    }
    *reinterpret_cast<float4*>(out.p + (long long)b * out.sb + (long long)c * out.sc + (long long)f * To + 4 * v) = make_float4(y[0], y[1], y[2], y[3]);
}

static bool tv_vec4(const TV& v) { return ((reinterpret_cast<uintptr_t>(v.p) & 15) == 0) && (v.sb & 3) == 0 && (v.sc & 3) == 0 && (v.T & 3) == 0; }

void launch_resample_down(const TV& x, const TV& out, cudaStream_t s) {
    if (tv_vec4(out) && tv_vec4(x) && out.T >= 8) {
        dim3 grid((unsigned)(((long long)out.F * (out.T / 4) + kThreads - 1) / kThreads), out.C, out.B);
        resample_down4_kernel<<<grid, kThreads, 0, s>>>(x, out);
    } else {
        dim3 grid((unsigned)(((long long)out.F * out.T + kThreads - 1) / kThreads), out.C, out.B);
        resample_down_kernel<<<grid, kThreads, 0, s>>>(x, out);
    }
    AID_COUNT_LAUNCH(1);
}
void launch_resample_up(const TV& x, const TV& out, cudaStream_t s) {
    if (tv_vec4(out) && x.T >= 4) {
        dim3 grid((unsigned)(((long long)out.F * (out.T / 4) + kThreads - 1) / kThreads), out.C, out.B);
        resample_up4_kernel<<<grid, kThreads, 0, s>>>(x, out);
    } else {
        dim3 grid((unsigned)(((long long)out.F * out.T + kThreads - 1) / kThreads), out.C, out.B);
        resample_up_kernel<<<grid, kThreads, 0, s>>>(x, out);
    }
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// One CTA per sigma.  emb[i] = relu(W2 relu(W1 relu(W0 [sin, cos](2*pi*s*f) + b0) + b1) + b2)   (64->128->256->256)
__global__ void __launch_bounds__(256)
embedding_kernel(const float* __restrict__ c_noise, const float* __restrict__ rff, const float* __restrict__ w0,
                 const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ emb) {
    __shared__ float e0[64], h1[128], h2[256];
    const int i = blockIdx.x, t = threadIdx.x;
    if (t < 32) {
        const float tp = 6.283185307179586f * c_noise[i];  // fp32(2*pi) * s, then * f: the reference's op order
        const float ang = tp * rff[t];
        e0[t] = sinf(ang);
        e0[32 + t] = cosf(ang);
    }
    __syncthreads();
    if (t < 128) {
        float a = 0.f;
        for (int j = 0; j < 64; ++j) a += w0[t * 64 + j] * e0[j];
        h1[t] = fmaxf(a + b0[t], 0.f);
    }
    __syncthreads();
    {
        float a = 0.f;
        for (int j = 0; j < 128; ++j) a += w1[t * 128 + j] * h1[j];
        h2[t] = fmaxf(a + b1[t], 0.f);
    }
    __syncthreads();
    {
        float a = 0.f;
        for (int j = 0; j < 256; ++j) a += w2[t * 256 + j] * h2[j];
        emb[i * 256 + t] = fmaxf(a + b2[t], 0.f);
    }
}

void launch_embedding(const float* c_noise, int n_sigma, const float* rff, const float* w0, const float* b0,
                      const float* w1, const float* b1, const float* w2, const float* b2, float* emb, cudaStream_t s) {
    embedding_kernel<<<n_sigma, 256, 0, s>>>(c_noise, rff, w0, b0, w1, b1, w2, b2, emb);
    AID_COUNT_LAUNCH(1);
}

// out[i][r] = bias[r] + sum_j W[r][j] * emb[i][j];  one warp per row.  grid: (ceil(total/8), n_sigma)
__global__ void __launch_bounds__(256)
mod_vectors_kernel(const float* __restrict__ emb, const float* __restrict__ W, const float* __restrict__ bias, int total,
                   float* __restrict__ out) {
    __shared__ float e[256];
    const int i = blockIdx.y;
    e[threadIdx.x] = emb[i * 256 + threadIdx.x];
    __syncthreads();
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (r >= total) return;
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) a += __ldg(W + (long long)r * 256 + q * 32 + l) * e[q * 32 + l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (l == 0) out[(long long)i * total + r] = a + bias[r];
}

void launch_mod_vectors(const float* emb, int n_sigma, const float* W, const float* bias, int total, float* out,
                        cudaStream_t s) {
    dim3 grid((total + 7) / 8, n_sigma);
    mod_vectors_kernel<<<grid, 256, 0, s>>>(emb, W, bias, total, out);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// x += scale * eps                                                                     sampler.py:214
__global__ void axpy_noise_kernel(float* __restrict__ x, const float* __restrict__ eps, float scale, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] += scale * eps[i];
}
void launch_axpy_noise(float* x, const float* eps, float scale, long long n, cudaStream_t s) {
    const int blocks = (int)min((long long)device_sm_count() * 8, (n + 255) / 256);
    axpy_noise_kernel<<<blocks, 256, 0, s>>>(x, eps, scale, n);
    AID_COUNT_LAUNCH(1);
}

// One fused sampler update over [B, L] (n = B*L elements).
//   xh  = mask ? mask*y + (1-mask)*xhat : xhat          (sampler.py:343 projection; mask indexed modulo L)
//   d   = (xin - xh) / sigma                            (= -sigma * score, sampler.py:147, 230)
//   mode 0: d_out = d,  x_out = xin + h*d               (Euler / Heun predictor, sampler.py:240, 251)
//   mode 1: x_out = xbase + h*(0.5*d_prev + 0.5*d)      (Heun corrector, sampler.py:247)
__global__ void edm_step_kernel(const float* __restrict__ xin, const float* __restrict__ xhat, const float* __restrict__ y,
                                const float* __restrict__ mask, long long mask_n, long long n, float sigma, float h,
                                int mode, const float* __restrict__ d_prev, const float* __restrict__ xbase,
                                float* __restrict__ d_out, float* __restrict__ x_out, const float* __restrict__ sigma_h_dev) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (sigma_h_dev) { sigma = sigma_h_dev[0]; h = sigma_h_dev[1]; }
    const float inv = 1.f / sigma;
    for (; i < n; i += stride) {
        float xh = xhat[i];
        if (mask) {
            const float m = mask[i % mask_n];
            xh = m * y[i] + (1.f - m) * xh;
        }
        const float d = (xin[i] - xh) * inv;
        if (mode == 0) {
            if (d_out) d_out[i] = d;
            x_out[i] = xin[i] + h * d;
        } else {
            x_out[i] = xbase[i] + h * (0.5f * d_prev[i] + 0.5f * d);
        }
    }
}

void launch_edm_step(const float* xin, const float* xhat, const float* y, const float* mask, long long mask_n, long long n,
                     float sigma, float h, int mode, const float* d_prev, const float* xbase, float* d_out, float* x_out,
                     cudaStream_t s, const float* sigma_h_dev) {
    const int blocks = (int)min((long long)device_sm_count() * 8, (n + 255) / 256);
    edm_step_kernel<<<blocks, 256, 0, s>>>(xin, xhat, y, mask, mask_n, n, sigma, h, mode, d_prev, xbase, d_out, x_out, sigma_h_dev);
    AID_COUNT_LAUNCH(1);
}

// ---------------------------------------------------------------------------------------------------
// Device-resident noise (edm.py:94, sampler.py:212 draw it with the CPU generator and copy 4*B*L bytes per step).
// Philox4x32-10 (Salmon et al., SC'11): counter = (i / 4, draw, global clip index, stream id), key = 64-bit seed; the four
// 32-bit outputs of one call become elements 4q .. 4q+3 of the clip through two Box-Muller pairs.  A clip's noise depends only
// on (seed, stream id, clip index, draw, element), never on the batch it is sampled in or on the number of ranks.
// oracle/philox_oracle.py restates it in numpy (integer stream bit-exact, normals to float32 rounding).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
// u in (0, 1): the top 24 bits, centred in their bin
__device__ __forceinline__ float philox_uniform(unsigned int r) { return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-8f; }
__device__ __forceinline__ void box_muller(unsigned int a, unsigned int b, float& n0, float& n1) {
    const float rad = sqrtf(-2.f * logf(philox_uniform(a)));
    float sn, cs;
    sincospif(2.f * philox_uniform(b), &sn, &cs);
    n0 = rad * cs; n1 = rad * sn;
}

// grid: (blocks over L/4 quads, n_clips)
__global__ void __launch_bounds__(256)
philox_normal_kernel(float* __restrict__ x, long long L, uint2 key, unsigned int stream_id, unsigned int clip0, unsigned int draw,
                     float scale, int accumulate, const float* __restrict__ scale_draw_dev) {
    if (scale_draw_dev) {
        scale = scale_draw_dev[0]; draw = __float_as_uint(scale_draw_dev[1]);
        stream_id = __float_as_uint(scale_draw_dev[2]); clip0 = __float_as_uint(scale_draw_dev[3]);
    }
    const unsigned int clip = clip0 + blockIdx.y;
    float* row = x + (long long)blockIdx.y * L;
    const long long nq = (L + 3) >> 2;
    const bool vec = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((unsigned int)q, draw, clip, stream_id), key);
        float n[4];
        box_muller(r.x, r.y, n[0], n[1]);
        box_muller(r.z, r.w, n[2], n[3]);
        if (vec) {
            float4* p4 = reinterpret_cast<float4*>(row) + q;
            float4 v = accumulate ? *p4 : make_float4(0.f, 0.f, 0.f, 0.f);
            v.x += scale * n[0]; v.y += scale * n[1]; v.z += scale * n[2]; v.w += scale * n[3];
            *p4 = v;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long i = 4 * q + j;
                if (i < L) row[i] = (accumulate ? row[i] : 0.f) + scale * n[j];
            }
        }
    }
}

void launch_philox_normal(float* x, int n_clips, long long L, unsigned long long seed, unsigned int stream_id, unsigned int clip0,
                          unsigned int draw, float scale, bool accumulate, const float* scale_draw_dev, cudaStream_t s) {
    if (n_clips <= 0 || L <= 0) return;
    const long long nq = (L + 3) / 4;
    const int bx = (int)max(1ll, min((nq + 255) / 256, (long long)(device_sm_count() * 8 + n_clips - 1) / n_clips));
    philox_normal_kernel<<<dim3(bx, n_clips), 256, 0, s>>>(x, L, make_uint2((unsigned int)(seed & 0xffffffffull), (unsigned int)(seed >> 32)),
                                                           stream_id, clip0, draw, scale, accumulate ? 1 : 0, scale_draw_dev);
    AID_COUNT_LAUNCH(1);
}

__global__ void sched_select_kernel(const float* __restrict__ table, int row, int* __restrict__ counter, float* __restrict__ cur) {
    const int k = *counter;
    if ((int)threadIdx.x < row) cur[threadIdx.x] = table[(long long)k * row + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) *counter = k + 1;
}
void launch_sched_select(const float* table, int row, int* counter, float* cur, cudaStream_t s) {
    sched_select_kernel<<<1, 64, 0, s>>>(table, row, counter, cur);
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
