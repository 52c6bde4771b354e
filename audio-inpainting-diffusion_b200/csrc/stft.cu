// Spectrogram-inpainting degradation and its projection (sm_100a), hand-written STFT -> mask -> inverse STFT.
//
// Replaces Sampler.apply_spectral_mask (sampler.py:271-290: zero-pad to a multiple of n_fft, torch.stft with a periodic Hann
// window, centre = True / reflect padding, multiply by a real [n_fft/2+1, frames] mask, torch.istft, crop) and the projection
// of the spectrogram mode, proj(x) = y + x - S(x) (sampler.py:361), which runs once per denoiser evaluation:
//
//   frames[b][t][j] = w[j] * Re IFFT_N( mask[min(k, N-k)][t] * FFT_N( w[.] * xp[b][t*hop - N/2 + .] ) )[j]
//   S(x)[b][n]      = sum_t frames[b][t][n + N/2 - t*hop] / sum_t w[n + N/2 - t*hop]^2
//
// xp = x zero-extended to Lp = L + (N - L % N) and reflected by N/2 at both ends.  One CTA transforms one frame in shared
// memory (radix-2, twiddles built per CTA with sincospif); the overlap-add is a gather (deterministic, no atomics) fused with
// the projection.  The work is ~0.1 % of a denoiser evaluation; the kernels are HBM / latency bound.
#include "common.cuh"

namespace aid {

static constexpr int ST_THREADS = 256;

__device__ __forceinline__ int st_brev(int i, int lg) { return (int)(__brev((unsigned)i) >> (32 - lg)); }
__device__ __forceinline__ float st_hann(int j, int N) { return 0.5f - 0.5f * cospif(2.f * (float)j / (float)N); }   // periodic Hann

// in-place radix-2 DIT FFT of n = 1 << lg points in shared memory, input in bit-reversed order; tw[k] = exp(-2 pi i k / n), k < n/2
__device__ __forceinline__ void st_fft(float2* arr, int lg, const float2* tw, bool inverse) {
    const int n = 1 << lg;
    for (int s = 1; s <= lg; ++s) {
        const int half = 1 << (s - 1);
        for (int j = threadIdx.x; j < (n >> 1); j += ST_THREADS) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> (s - 1)) << s) + pos, i1 = i0 + half;
            float2 w = tw[pos << (lg - s)];
            if (inverse) w.y = -w.y;
            const float2 a = arr[i1], u = arr[i0];
            const float2 t = make_float2(w.x * a.x - w.y * a.y, w.x * a.y + w.y * a.x);
            arr[i0] = make_float2(u.x + t.x, u.y + t.y);
            arr[i1] = make_float2(u.x - t.x, u.y - t.y);
        }
        __syncthreads();
    }
}

// grid: (n_frames, B).  shared: N + N/2 float2
__global__ void __launch_bounds__(ST_THREADS)
stft_mask_frames_kernel(const float* __restrict__ x, int L, int Lp, int lgN, int hop, int n_frames, const float* __restrict__ mask,
                        float* __restrict__ frames) {
    extern __shared__ float2 st_sm[];
    const int N = 1 << lgN;
    float2* arr = st_sm;
    float2* tw = st_sm + N;
    const int t = blockIdx.x, b = blockIdx.y;
    for (int k = threadIdx.x; k < (N >> 1); k += ST_THREADS) {
        float s, c;
        sincospif(-2.f * (float)k / (float)N, &s, &c);
        tw[k] = make_float2(c, s);
    }
    const float* xb = x + (long long)b * L;
    for (int j = threadIdx.x; j < N; j += ST_THREADS) {
        int i = t * hop - (N >> 1) + j;
        if (i < 0) i = -i;
        else if (i >= Lp) i = 2 * (Lp - 1) - i;
        const float v = i < L ? __ldg(xb + i) : 0.f;
        arr[st_brev(j, lgN)] = make_float2(st_hann(j, N) * v, 0.f);
    }
    __syncthreads();
    st_fft(arr, lgN, tw, false);
    // real mask on both halves of the spectrum, and the bit-reversal permutation the inverse transform wants
    for (int k = threadIdx.x; k < N; k += ST_THREADS) {
        const int r = st_brev(k, lgN);
        if (k <= r) {
            const float mk = __ldg(mask + (long long)min(k, N - k) * n_frames + t), mr = __ldg(mask + (long long)min(r, N - r) * n_frames + t);
            const float2 a = arr[k], c = arr[r];
            arr[k] = make_float2(c.x * mr, c.y * mr);
            arr[r] = make_float2(a.x * mk, a.y * mk);
        }
    }
    __syncthreads();
    st_fft(arr, lgN, tw, true);
    float* fo = frames + ((long long)b * n_frames + t) * N;
    const float inv = 1.f / (float)N;
    for (int j = threadIdx.x; j < N; j += ST_THREADS) fo[j] = st_hann(j, N) * arr[j].x * inv;
}

// out[b][n] = y ? y + x - S : S        grid: ceil(B * L / 256)
__global__ void __launch_bounds__(ST_THREADS)
istft_project_kernel(const float* __restrict__ frames, const float* __restrict__ x, const float* __restrict__ y, long long total, int L,
                     int lgN, int hop, int n_frames, float* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * ST_THREADS + threadIdx.x;
    if (idx >= total) return;
    const int N = 1 << lgN;
    const int b = (int)(idx / L), n = (int)(idx - (long long)b * L);
    const int p = n + (N >> 1);                               // position in the centred (reflect-padded) frame grid
    const int t_lo = p >= N ? (p - N) / hop + 1 : 0, t_hi = min(p / hop, n_frames - 1);
    float acc = 0.f, env = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
        const int j = p - t * hop;
        const float w = st_hann(j, N);
        acc += __ldg(frames + ((long long)b * n_frames + t) * N + j);
        env = fmaf(w, w, env);
    }
    const float s = acc / env;
    out[idx] = y ? y[idx] + x[idx] - s : s;
}

void launch_spectral_mask(const float* x, const float* y, const float* mask, int B, int L, int n_fft, int hop, int n_frames,
                          float* frames, float* out, cudaStream_t s) {
    int lgN = 0;
    while ((1 << lgN) < n_fft) ++lgN;
    const int Lp = L + (n_fft - L % n_fft);
    const size_t smem = (size_t)(n_fft + n_fft / 2) * sizeof(float2);
    if (smem > 48 * 1024)
        AID_CUDA_CHECK(cudaFuncSetAttribute(stft_mask_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_mask_frames_kernel<<<dim3((unsigned)n_frames, (unsigned)B), ST_THREADS, smem, s>>>(x, L, Lp, lgN, hop, n_frames, mask, frames);
    const long long total = (long long)B * L;
    istft_project_kernel<<<(unsigned)((total + ST_THREADS - 1) / ST_THREADS), ST_THREADS, 0, s>>>(frames, x, y, total, L, lgN, hop, n_frames, out);
    AID_COUNT_LAUNCH(2);
}

}  // namespace aid
