// Hand-written FFTs and the constant-Q (NSGT, "oct" mode) analysis / synthesis kernels (sm_100a).
//
// Replaces cqt_nsgt_pytorch.CQT_nsgt.fwd / .bwd / .apply_hpf_DC as called from unet.py:743, unet.py:841 and
// sampler.py:63,123.  That package is not part of the reference tree; the definition implemented here is
// the one written down in DESIGN.md (section "CQT") and restated on the CPU in oracle/cqt_oracle.py.
//
//   analysis : X = FFT_L(x);  coef_k = IFFT_M( fold( X[(c_k+m) mod L] * w_k(m) ) )        per band k
//   synthesis: fr[n] = sum_k FFT_M(coef_k)[(n-c_k) mod M] * dual_k(n-c_k);  x = irFFT_L(fr)
//
// Length-L transforms use the four-step split L = N1*N2 with both factors done in shared memory;
// band transforms (M <= 4096) are single shared-memory radix-2 FFTs.  All twiddles come from one
// table W_L^k built in double precision on the host.
#include "common.cuh"

namespace aid {

static constexpr int FNT = 256;
static constexpr int CW = 8;  // columns (step 1) / rows (step 3) per CTA

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// `nfft` independent radix-2 DIT FFTs of size 1<<lg stored back to back in shared memory, inputs already in
// bit-reversed order.  tw = W_L^k table, lgL = log2(L).  All threads of the CTA must call this.
__device__ __forceinline__ void smem_fft(float2* arr, int lg, int nfft, const float2* __restrict__ tw, int lgL,
                                         bool inverse) {
    const int n = 1 << lg;
    const int total = nfft * (n >> 1);
    for (int s = 1; s <= lg; ++s) {
        const int half = 1 << (s - 1);
        for (int j = threadIdx.x; j < total; j += blockDim.x) {
            const int fi = j >> (lg - 1), jj = j & ((n >> 1) - 1);
            const int pos = jj & (half - 1);
            const int i0 = fi * n + ((jj >> (s - 1)) << s) + pos, i1 = i0 + half;
            float2 w = __ldg(tw + ((long long)pos << (lgL - s)));
            if (inverse) w.y = -w.y;
            const float2 t = cmul(w, arr[i1]);
            const float2 u = arr[i0];
            arr[i0] = make_float2(u.x + t.x, u.y + t.y);
            arr[i1] = make_float2(u.x - t.x, u.y - t.y);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int brev(int i, int lg) { return lg == 0 ? 0 : (int)(__brev((unsigned)i) >> (32 - lg)); }

// ---- four-step, step 1+2: FFT over n1 for CW adjacent columns n2, then twiddle W_L^(n2*k1) -----------
// grid: (N2/CW, B)
__global__ void __launch_bounds__(FNT)
fft_cols_kernel(int lg1, int lg2, const float2* __restrict__ tw, bool inverse, const float* __restrict__ in_real,
                long long in_stride, float in_scale, const float2* __restrict__ in_cplx, float2* __restrict__ tmp,
                const float* __restrict__ dscal) {
    extern __shared__ float2 arr[];
    if (dscal) in_scale *= dscal[0];
    const int N1 = 1 << lg1, N2 = 1 << lg2, lgL = lg1 + lg2;
    const long long L = (long long)N1 * N2;
    const int b = blockIdx.y, n20 = blockIdx.x * CW;
    for (int e = threadIdx.x; e < N1 * CW; e += FNT) {
        const int n1 = e / CW, c = e % CW;
        const long long idx = (long long)n1 * N2 + n20 + c;
        float2 v;
        if (in_real) v = make_float2(in_real[(long long)b * in_stride + idx] * in_scale, 0.f);
        else v = in_cplx[(long long)b * L + idx];
        arr[c * N1 + brev(n1, lg1)] = v;
    }
    __syncthreads();
    smem_fft(arr, lg1, CW, tw, lgL, inverse);
    for (int e = threadIdx.x; e < N1 * CW; e += FNT) {
        const int k1 = e / CW, c = e % CW;
        float2 w = __ldg(tw + (long long)(n20 + c) * k1);  // (n2*k1) < L
        if (inverse) w.y = -w.y;
        tmp[(long long)b * L + (long long)k1 * N2 + n20 + c] = cmul(arr[c * N1 + k1], w);
    }
}

// ---- four-step, step 3: FFT over n2 for CW adjacent rows k1; X[k1 + N1*k2] ----------------------------
// grid: (N1/CW, B)
__global__ void __launch_bounds__(FNT)
fft_rows_kernel(int lg1, int lg2, const float2* __restrict__ tw, bool inverse, const float2* __restrict__ tmp,
                float2* __restrict__ out_cplx, float* __restrict__ out_real, long long out_stride, float out_scale,
                const float* __restrict__ skip, long long skip_stride, float skip_scale, const float* __restrict__ dscal) {
    extern __shared__ float2 arr[];
    if (dscal) { out_scale *= dscal[1]; skip_scale = dscal[2]; }
    const int N1 = 1 << lg1, N2 = 1 << lg2, lgL = lg1 + lg2;
    const long long L = (long long)N1 * N2;
    const int b = blockIdx.y, k10 = blockIdx.x * CW;
    for (int e = threadIdx.x; e < N2 * CW; e += FNT) {
        const int r = e >> lg2, n2 = e & (N2 - 1);
        arr[r * N2 + brev(n2, lg2)] = tmp[(long long)b * L + (long long)(k10 + r) * N2 + n2];
    }
    __syncthreads();
    smem_fft(arr, lg2, CW, tw, lgL, inverse);
    for (int e = threadIdx.x; e < N2 * CW; e += FNT) {
        const int k2 = e / CW, r = e % CW;
        const long long k = (long long)(k10 + r) + (long long)N1 * k2;
        const float2 v = arr[r * N2 + k2];
        if (out_cplx) out_cplx[(long long)b * L + k] = v;
        if (out_real) {
            float o = out_scale * v.x;
            if (skip) o += skip_scale * skip[(long long)b * skip_stride + k];
            out_real[(long long)b * out_stride + k] = o;
        }
    }
}

static void fft_pow2(const FftPlan& plan, int B, bool inverse, const float* in_real, long long in_stride, float in_scale,
                     const float2* in_cplx, float2* tmp, float2* out_cplx, float* out_real, long long out_stride, float out_scale,
                     const float* skip, long long skip_stride, float skip_scale, cudaStream_t s) {
    int lg1 = 0, lg2 = 0;
    while ((1 << lg1) < plan.N1) ++lg1;
    while ((1 << lg2) < plan.N2) ++lg2;
    const size_t sm1 = (size_t)plan.N1 * CW * sizeof(float2), sm2 = (size_t)plan.N2 * CW * sizeof(float2);
    static SmemConfig c1, c2;
    ensure_dyn_smem(fft_cols_kernel, sm1, c1);
    ensure_dyn_smem(fft_rows_kernel, sm2, c2);
    fft_cols_kernel<<<dim3(plan.N2 / CW, B), FNT, sm1, s>>>(lg1, lg2, plan.tw, inverse, in_real, in_stride, in_scale, in_cplx, tmp,
                                                            in_real ? plan.dscal : nullptr);
    AID_COUNT_LAUNCH(1);
    fft_rows_kernel<<<dim3(plan.N1 / CW, B), FNT, sm2, s>>>(lg1, lg2, plan.tw, inverse, tmp, out_cplx, out_real, out_stride,
                                                            out_scale, skip, skip_stride, skip_scale, out_real ? plan.dscal : nullptr);
    AID_COUNT_LAUNCH(1);
}

// ---- Bluestein (chirp-z) for lengths that are not a power of two (the reference's trained audio_len = 184184) --------
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w)[k-n],  w[n] = exp(-i*pi*n^2/L): one circular convolution of power-of-two size M >= 2L-1.
//   The inverse transform is conj(DFT(conj(x))).
__global__ void bluestein_pre_kernel(int L, int M, bool inverse, const float* __restrict__ in_real, long long in_stride, float in_scale,
                                     const float2* __restrict__ in_cplx, const float2* __restrict__ chirp, float2* __restrict__ A,
                                     const float* __restrict__ dscal) {
    const int b = blockIdx.y;
    if (dscal) in_scale *= dscal[0];
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x) {
        float2 v = make_float2(0.f, 0.f);
        if (n < L) {
            float2 x = in_real ? make_float2(in_real[(long long)b * in_stride + n] * in_scale, 0.f) : in_cplx[(long long)b * L + n];
            if (inverse) x.y = -x.y;
            v = cmul(x, __ldg(chirp + n));
        }
        A[(long long)b * M + n] = v;
    }
}
__global__ void bluestein_mul_kernel(int M, const float2* __restrict__ bfilt, float2* __restrict__ A) {
    const int b = blockIdx.y;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < M; n += gridDim.x * blockDim.x)
        A[(long long)b * M + n] = cmul(A[(long long)b * M + n], __ldg(bfilt + n));
}
__global__ void bluestein_post_kernel(int L, int M, bool inverse, const float2* __restrict__ A, const float2* __restrict__ chirp,
                                      float2* __restrict__ out_cplx, float* __restrict__ out_real, long long out_stride, float out_scale,
                                      const float* __restrict__ skip, long long skip_stride, float skip_scale,
                                      const float* __restrict__ dscal) {
    const int b = blockIdx.y;
    const float inv = 1.f / (float)M;
    if (dscal) { out_scale *= dscal[1]; skip_scale = dscal[2]; }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L; k += gridDim.x * blockDim.x) {
        float2 v = cmul(A[(long long)b * M + k], __ldg(chirp + k));
        v.x *= inv; v.y *= inv;
        if (inverse) v.y = -v.y;
        if (out_cplx) out_cplx[(long long)b * L + k] = v;
        if (out_real) {
            float o = out_scale * v.x;
            if (skip) o += skip_scale * skip[(long long)b * skip_stride + k];
            out_real[(long long)b * out_stride + k] = o;
        }
    }
}

void launch_fft_big(const FftPlan& plan, int B, bool inverse, const float* in_real, long long in_stride, float in_scale,
                    const float2* in_cplx, float2* tmp, float2* scratch, float2* out_cplx, float* out_real, long long out_stride,
                    float out_scale, const float* skip, long long skip_stride, float skip_scale, cudaStream_t s) {
    if (plan.M == plan.L) {
        fft_pow2(plan, B, inverse, in_real, in_stride, in_scale, in_cplx, tmp, out_cplx, out_real, out_stride, out_scale, skip,
                 skip_stride, skip_scale, s);
        return;
    }
    const dim3 grid(device_sm_count() * 4, B);
    FftPlan inner = plan; inner.dscal = nullptr;     // the device scalars apply to the outer real input / output only
    bluestein_pre_kernel<<<grid, 256, 0, s>>>(plan.L, plan.M, inverse, in_real, in_stride, in_scale, in_cplx, plan.chirp, scratch,
                                              in_real ? plan.dscal : nullptr);
    fft_pow2(inner, B, false, nullptr, 0, 0.f, scratch, tmp, scratch, nullptr, 0, 0.f, nullptr, 0, 0.f, s);
    bluestein_mul_kernel<<<grid, 256, 0, s>>>(plan.M, plan.bfilt, scratch);
    fft_pow2(inner, B, true, nullptr, 0, 0.f, scratch, tmp, scratch, nullptr, 0, 0.f, nullptr, 0, 0.f, s);
    bluestein_post_kernel<<<grid, 256, 0, s>>>(plan.L, plan.M, inverse, scratch, plan.chirp, out_cplx, out_real, out_stride, out_scale,
                                               skip, skip_stride, skip_scale, out_real ? plan.dscal : nullptr);
    AID_COUNT_LAUNCH(3);
}

// ---- CQT analysis: one CTA per (band of the octave, clip) ---------------------------------------------
__global__ void __launch_bounds__(FNT)
cqt_analysis_kernel(CqtTables t, const float2* __restrict__ tw, int lgL, int oct, int lgM, const float2* __restrict__ spec,
                    TV C) {
    extern __shared__ float2 arr[];
    const int M = 1 << lgM;
    const int kk = blockIdx.x, b = blockIdx.y;
    const int k = oct * t.bins + kk;
    const int lg = t.Lg[k], cen = t.centre[k], wo = t.woff[k];
    const int r = lg - lg / 2, l = lg / 2;
    const float2* X = spec + (long long)b * t.L;
    for (int j = threadIdx.x; j < M; j += FNT) {
        float2 v = make_float2(0.f, 0.f);
        int m = 0; bool on = false;
        if (j < r) { m = j; on = true; }
        else if (j >= M - l) { m = j - M; on = true; }
        if (on) {
            int bin = cen + m;
            if (bin < 0) bin += t.L;
            if (bin >= t.L) bin -= t.L;
            const float w = __ldg(t.win + wo + (m >= 0 ? m : lg + m));
            const float2 x = X[bin];
            v = make_float2(x.x * w, x.y * w);
        }
        arr[brev(j, lgM)] = v;
    }
    __syncthreads();
    smem_fft(arr, lgM, 1, tw, lgL, /*inverse=*/true);
    const float inv = 1.f / (float)M;
    float* re = C.p + (long long)b * C.sb + (long long)kk * C.T;
    float* im = re + C.sc;
    for (int n = threadIdx.x; n < M; n += FNT) {
        re[n] = arr[n].x * inv;
        im[n] = arr[n].y * inv;
    }
}

void launch_cqt_analysis_oct(const CqtTables& t, const FftPlan& fp, int oct, const float2* spec, const TV& C, cudaStream_t s) {
    const int M = t.M[oct];
    int lgM = 0;
    while ((1 << lgM) < M) ++lgM;
    const int lgL = fp.lgM;  // log2 of the twiddle table size
    const size_t sm = (size_t)M * sizeof(float2);
    static size_t cfg = 0;
    if (sm > cfg) { AID_CUDA_CHECK(cudaFuncSetAttribute(cqt_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); cfg = sm; }
    cqt_analysis_kernel<<<dim3(t.bins, C.B), FNT, sm, s>>>(t, fp.tw, lgL, oct, lgM, spec, C);
    AID_COUNT_LAUNCH(1);
}

// ---- CQT synthesis, step 1: Y_k = FFT_M(coef_k) ---------------------------------------------------------
__global__ void __launch_bounds__(FNT)
cqt_synth_kernel(CqtTables t, const float2* __restrict__ tw, int lgL, int oct, int lgM, TV C, float2* __restrict__ Y) {
    extern __shared__ float2 arr[];
    const int M = 1 << lgM;
    const int kk = blockIdx.x, b = blockIdx.y;
    const float* re = C.p + (long long)b * C.sb + (long long)kk * C.T;
    const float* im = re + C.sc;
    for (int n = threadIdx.x; n < M; n += FNT) arr[brev(n, lgM)] = make_float2(re[n], im[n]);
    __syncthreads();
    smem_fft(arr, lgM, 1, tw, lgL, /*inverse=*/false);
    float2* y = Y + (long long)b * t.ytotal + t.yoff[oct] + (long long)kk * M;
    for (int n = threadIdx.x; n < M; n += FNT) y[n] = arr[n];
}

void launch_cqt_synth_oct(const CqtTables& t, const FftPlan& fp, int oct, const TV& C, float2* Y, cudaStream_t s) {
    const int M = t.M[oct];
    int lgM = 0;
    while ((1 << lgM) < M) ++lgM;
    const int lgL = fp.lgM;  // log2 of the twiddle table size
    const size_t sm = (size_t)M * sizeof(float2);
    static size_t cfg = 0;
    if (sm > cfg) { AID_CUDA_CHECK(cudaFuncSetAttribute(cqt_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); cfg = sm; }
    cqt_synth_kernel<<<dim3(t.bins, C.B), FNT, sm, s>>>(t, fp.tw, lgL, oct, lgM, C, Y);
    AID_COUNT_LAUNCH(1);
}

// ---- CQT synthesis, step 2: deterministic overlap-add by gathering, Hermitian completion ----------------
// grid: (ceil((L/2+1)/256), B)
__global__ void __launch_bounds__(FNT) cqt_gather_kernel(CqtTables t, const float2* __restrict__ Y, float2* __restrict__ fr) {
    const int n = blockIdx.x * FNT + threadIdx.x, b = blockIdx.y;
    const int half = t.L >> 1;
    if (n > half) return;
    const float2* y = Y + (long long)b * t.ytotal;
    float2 acc = make_float2(0.f, 0.f);
    const int lo = t.klo[n], hi = t.khi[n];
    for (int k = lo; k <= hi; ++k) {
        const int lg = t.Lg[k], m = n - t.centre[k];
        if (m < -(lg / 2) || m >= lg - lg / 2) continue;
        const int o = k / t.bins, M = t.M[o];
        const float d = __ldg(t.dual + t.woff[k] + (m >= 0 ? m : lg + m));
        const float2 v = y[t.yoff[o] + (long long)(k - o * t.bins) * M + (m >= 0 ? m : M + m)];
        acc.x += v.x * d; acc.y += v.y * d;
    }
    float2* f = fr + (long long)b * t.L;
    if (n == 0 || n == half) { f[n] = make_float2(acc.x, 0.f); }  // irfft ignores these imaginary parts
    else { f[n] = acc; f[t.L - n] = make_float2(acc.x, -acc.y); }
}

void launch_cqt_synth_gather(const CqtTables& t, int B, const float2* Y, float2* fr, cudaStream_t s) {
    cqt_gather_kernel<<<dim3((t.L / 2 + 1 + FNT - 1) / FNT, B), FNT, 0, s>>>(t, Y, fr);
    AID_COUNT_LAUNCH(1);
}

__global__ void spec_mul_real_kernel(long long n, int L, float2* __restrict__ spec, const float* __restrict__ h) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const float g = __ldg(h + (i % L));
        float2 v = spec[i];
        spec[i] = make_float2(v.x * g, v.y * g);
    }
}
void launch_spec_mul_real(int B, int L, float2* spec, const float* h, cudaStream_t s) {
    const long long n = (long long)B * L;
    const int blocks = (int)min((long long)device_sm_count() * 8, (n + 255) / 256);
    spec_mul_real_kernel<<<blocks, 256, 0, s>>>(n, L, spec, h);
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
