// Shared device/host types for the B200 denoiser path.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace aid {

// A [B, C, F, T] fp32 activation view.  T is contiguous and the (F, T) plane of one channel is
// contiguous (F-slices of a wider buffer keep that property), so only batch and channel strides are free.
// `stats` (optional) points at this tensor's group-norm accumulators: [B][8][2] doubles = (sum, sum of squares)
// over the 8 channel groups of the *whole* logical tensor the view belongs to (unet.py:147-163).
struct TV {
    float* p = nullptr;
    int B = 0, C = 0, F = 0, T = 0;
    long long sb = 0, sc = 0;  // element strides
    double* stats = nullptr;
};

static inline TV make_tv(float* p, int B, int C, int F, int T) {
    TV v; v.p = p; v.B = B; v.C = C; v.F = F; v.T = T; v.sc = (long long)F * T; v.sb = v.sc * C; return v;
}
// channels-last [B][F][T][C] tensor described by a TV (sb = F*T*C, channel stride 1; only conv_tc2 / gn_act_tc2_cl take these)
static inline TV make_tv_cl(float* p, int B, int C, int F, int T) {
    TV v; v.p = p; v.B = B; v.C = C; v.F = F; v.T = T; v.sc = 1; v.sb = (long long)C * F * T; return v;
}
// rows [f0, f0+nf) of v
static inline TV slice_f(const TV& v, int f0, int nf) {
    TV r = v; r.p = v.p + (long long)f0 * v.T; r.F = nf; r.stats = nullptr; return r;
}
// channels [c0, c0+nc) of v
static inline TV slice_c(const TV& v, int c0, int nc) {
    TV r = v; r.p = v.p + (long long)c0 * v.sc; r.C = nc; r.stats = nullptr; return r;
}

struct ConvEpilogue {
    // out = alpha * (acc * gate[c] + R) + beta * R2        (unet.py:482, 470, 491, 794, 817)
    const float* gate = nullptr;  // per-channel gate vector (adaLN gate Linear output), or null => 1
    long long gate_bstride = 0;   // 0 when sigma is shared by the batch
    TV R;                         // R.p == nullptr => no residual
    TV R2;                        // R2.p == nullptr => none
    float alpha = 1.f, beta = 0.f;
    double* stats = nullptr;      // accumulate (sum, sumsq) of `out` per (b, group of Cout/8 channels)
    // conv_mode 2 only: R / out are channels-last [B][F][T][C] tensors (TV with sb = F*T*C; see make_tv_cl)
    bool R_cl = false, out_cl = false;
};

#define AID_CUDA_CHECK(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) throw aid::CudaError(_e, #expr, __FILE__, __LINE__);   \
    } while (0)

struct CudaError {
    cudaError_t code; const char* expr; const char* file; int line;
    CudaError(cudaError_t c, const char* e, const char* f, int l) : code(c), expr(e), file(f), line(l) {}
};

extern unsigned long long g_launch_count;  // kernels launched by this library (bench.py's gpu_launches)
#define AID_COUNT_LAUNCH(n) (aid::g_launch_count += (n))

// Per-device bookkeeping.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, so the size a
// kernel was configured with is remembered per device (a second handle on another GPU of the same process configures again).
static constexpr int kMaxDevices = 64;
struct SmemConfig { size_t bytes[kMaxDevices] = {0}; };
template <class Kernel>
inline void ensure_dyn_smem(Kernel kernel, size_t bytes, SmemConfig& cfg, size_t preset = 0) {
    int dev = 0;
    AID_CUDA_CHECK(cudaGetDevice(&dev));
    const int slot = dev < kMaxDevices ? dev : kMaxDevices - 1;
    if (dev < kMaxDevices && bytes <= (cfg.bytes[slot] ? cfg.bytes[slot] : preset)) return;
    AID_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cfg.bytes[slot] = bytes;
}
// multiprocessor count of the current device (cached per device); launch heuristics size their grids from it
int device_sm_count();
// RAII: make `device` current, restore the caller's device on scope exit (torch's current device is left alone)
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        AID_CUDA_CHECK(cudaGetDevice(&prev));
        if (prev != device) AID_CUDA_CHECK(cudaSetDevice(device)); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- launchers (defined in the .cu files) ---------------------------------------------------------
void launch_group_stats(const TV& x, double* stats, cudaStream_t s);
void launch_gn_act(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                   long long affine_bstride, bool gelu, const TV& out, cudaStream_t s);
void launch_combine(const TV& a, const TV& b, float alpha, float beta, const TV& out, double* stats, cudaStream_t s);
void launch_resample_down(const TV& x, const TV& out, cudaStream_t s);
void launch_resample_up(const TV& x, const TV& out, cudaStream_t s);
void launch_conv_simt(const TV& a, const float* wp, int KF, int KT, int dil, const TV& out, const ConvEpilogue& ep,
                      cudaStream_t s);
// thin convolutions (conv_thin.cu); returns false when the shape is not covered
bool launch_conv_thin(const TV& a, const float* wp, int KF, int KT, int dil, const TV& out, const ConvEpilogue& ep, cudaStream_t s);
// attention proj_in (N -> 8, 1x1) reading the un-normalised tensor: the group norm (no activation) is folded into per-clip weights
bool launch_conv_thin_out_normed(const TV& a, const double* stats, long long n_per_group, const float* gamma, const float* affine, long long affine_bstride,
                                 const float* wp, const TV& out, cudaStream_t s);
void launch_attention(const TV& h, const float* qk, const TV& out, cudaStream_t s);
// tcgen05 version (attention.cu): Q K^T with split-fp16 operands, softmax from TMEM, P V with fp16 operands
bool attention_tc_supported(int F, int T);
void launch_attention_tc(const TV& h, const float* qk, const TV& out, cudaStream_t s);
void launch_embedding(const float* c_noise, int n_sigma, const float* rff, const float* w0, const float* b0,
                      const float* w1, const float* b1, const float* w2, const float* b2, float* emb, cudaStream_t s);
void launch_mod_vectors(const float* emb, int n_sigma, const float* W, const float* bias, int total, float* out,
                        cudaStream_t s);

// tcgen05 path for the dilated 5x3 convolutions (conv_tc.cu).  Operands are split-fp16 planar: [B][C/8][F][T+2][8].

bool conv_tc_supported(int Cin, int Cout, int KF, int KT);
void launch_pack_weight_tc(const float* w, __half* wp, int Cout, int Cin, int KF, int KT, int parts, cudaStream_t s);
int tc_pad_rows(int T, int KF, int dil);  // zero rows needed above/below each operand plane (0 when T % 128 == 0)
void launch_gn_act_tc(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                      long long affine_bstride, bool gelu, int PF, __half* a_hi, __half* a_lo, cudaStream_t s);
void launch_to_planar_tc(const TV& x, int PF, __half* a_hi, __half* a_lo, cudaStream_t s);
void launch_conv_tc(const __half* a_hi, const __half* a_lo, int PF, const __half* wp, int B, int Cin, int F, int T, int KF, int KT, int dil,
                    const TV& out, const ConvEpilogue& ep, int num_sms, cudaStream_t s);

// second-generation tcgen05 path (conv_tc2.cu, conv_mode 2): single fp16 operands, channels-last [B][ceil(C/64)][F+2PF][T+2][64]
size_t tc2_weight_halves(int Cout, int Cin, int KF, int KT);
size_t tc2_act_halves(int B, int C, int F, int T, int PF);
void tc2_read_profile(unsigned long long* out16);
void launch_pack_weight_tc2(const float* w, __half* wp, int Cout, int Cin, int KF, int KT, cudaStream_t s, unsigned long long* sat = nullptr);
// stft.cu: out = y ? y + x - S(x) : S(x), S = crop(istft(mask * stft(zero-pad(x))))  (sampler.py:271-290, 361); frames: [B][n_frames][n_fft] scratch
void launch_spectral_mask(const float* x, const float* y, const float* mask, int B, int L, int n_fft, int hop, int n_frames,
                          float* frames, float* out, cudaStream_t s);
void launch_gn_act_tc2(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                       long long affine_bstride, bool gelu, int PF, __half* a, cudaStream_t s, unsigned long long* sat = nullptr);
void launch_to_planar_tc2(const TV& x, int PF, __half* a, cudaStream_t s, unsigned long long* sat = nullptr);
void launch_gn_act_tc2_cl(const float* x_cl, int B, int C, int F, int T, const double* stats, long long n_per_group, const float* gamma,
                          const float* affine, long long affine_bstride, bool gelu, int PF, __half* a, cudaStream_t s);
void launch_conv_tc2(const __half* a, int PF, const __half* wp, int B, int Cin, int F, int T, int KF, int KT, int dil,
                     const TV& out, const ConvEpilogue& ep, int num_sms, cudaStream_t s);

// fused dilated residual layer (conv_comb.cu): normalise + modulate + GELU + operand conversion inside the convolution kernel;
// ep.R must be x itself, out must not overlap x
// operand of [upsample2x(up) | x] without the fp32 copy of the upsampled half (conv_tc2.cu)
bool to_planar_tc2_up_supported(const TV& up, const TV& x);
void launch_to_planar_tc2_up(const TV& up, const TV& x, int PF, __half* a, cudaStream_t s);
// fused init block of an encoder level (conv_init.cu): proj_in (2 -> N) + one gated 1x1 residual layer + res_conv, one pass over the input
bool init_block_supported(int N, int T);
size_t init_block_scratch_floats(int B, int N);
void launch_init_block(const TV& x2, const float* w_in, const float* w_res, const __half* wH, const float* gamma, const float* affine,
                       long long affine_bstride, const float* gate, long long gate_bstride, const TV& out, double* stats_out, float* scratch,
                       int num_sms, cudaStream_t s);
// fused out block of the decoder / bottleneck (out_block.cu): one gated 1x1 layer + proj_out / res_conv N -> 2 collapsed into one pass over x
bool out_block_supported(const TV& x, const TV& out, const TV& accum);
size_t out_block_scratch_floats(int B, int N);
void launch_out_block(const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine, long long affine_bstride,
                      const float* gate, long long gate_bstride, const float* hw, const float* pw, const float* rw, const TV& out, const TV& accum,
                      float* scratch, cudaStream_t s);
bool conv_comb_supported(int C, int F, int T, int dil);
bool conv_comb_worthwhile(int B, int T, int dil, int num_sms);   // enough combs to fill the device
// 96 channels: the fused kernel has its own weight packing (64-channel group + 32-channel group per tap); 64 channels: launch_pack_weight_tc2's
size_t comb_weight_halves(int C);
void launch_pack_weight_comb(const float* w, __half* wp, int C, cudaStream_t s);
void launch_conv_comb(const TV& x, const double* stats_in, long long n_per_group, const float* gamma, const float* affine, long long affine_bstride,
                      const __half* wp, int dil, const TV& out, const ConvEpilogue& ep, int num_sms, cudaStream_t s);

// FFT / CQT
struct FftPlan {
    int L = 0;                      // transform length (even)
    int M = 0, N1 = 0, N2 = 0, lgM = 0;  // power-of-two engine: M == L if L is a power of two, else next_pow2(2L) (Bluestein)
    const float2* tw = nullptr;     // W_M^k = exp(-2*pi*i*k/M), k in [0, M)
    const float2* chirp = nullptr;  // Bluestein only: exp(-i*pi*n^2/L), n in [0, L)
    const float2* bfilt = nullptr;  // Bluestein only: FFT_M of the wrapped conj(chirp) filter
    // optional device scalars [in, out, skip] (CUDA-graph replay of the sampler: the preconditioning of the current sigma lives in
    // device memory): real input *= dscal[0], out_scale *= dscal[1], skip_scale = dscal[2] (a skip pointer must still be given)
    const float* dscal = nullptr;
};
// batched length-L complex DFT (four-step split M = N1*N2 in shared memory; Bluestein around it when L is not a power of two).
//  in_real != null : input is real [B][in_stride], scaled by in_scale;   in_cplx != null : input complex [B][L]
//  out_cplx != null: complex output [B][L] (unscaled)
//  out_real != null: out_real[b][n] = out_scale * Re(result) + skip_scale * skip[b][n]   (skip may be null)
//  tmp: [B][M] complex scratch; scratch: second [B][M] buffer, only needed when M != L.
void launch_fft_big(const FftPlan& plan, int B, bool inverse, const float* in_real, long long in_stride, float in_scale,
                    const float2* in_cplx, float2* tmp, float2* scratch, float2* out_cplx, float* out_real, long long out_stride,
                    float out_scale, const float* skip, long long skip_stride, float skip_scale, cudaStream_t s);

struct CqtTables {
    int L = 0, K = 0, bins = 0, nocts = 0;
    const int* centre = nullptr;   // [K] centre bin of band k (bands 1..K of the plan)
    const int* Lg = nullptr;       // [K] window length
    const int* woff = nullptr;     // [K] offset of band k's window in win / dual
    const float* win = nullptr;    // analysis windows, "peak at index 0" order
    const float* dual = nullptr;   // synthesis windows * M_o / D
    const int* klo = nullptr;      // [L/2+1] first band covering bin n (or K if none)
    const int* khi = nullptr;      // [L/2+1] last band covering bin n (or -1)
    const float* hhpf = nullptr;   // [L]
    int M[16] = {0};               // coefficients per band for octave o
    long long yoff[16] = {0};      // offset of octave o in the per-clip synthesis scratch (complex elements)
    long long ytotal = 0;          // complex elements per clip in the synthesis scratch
};
// spec [B][L] complex -> one octave's coefficients, written as C[b, 0/1, band_in_oct, frame] (re, im channels)
void launch_cqt_analysis_oct(const CqtTables& t, const FftPlan& fp, int oct, const float2* spec, const TV& C, cudaStream_t s);
// C (re, im channels) of one octave -> FFT_M of every band into scratch Y [B][ytotal]
void launch_cqt_synth_oct(const CqtTables& t, const FftPlan& fp, int oct, const TV& C, float2* Y, cudaStream_t s);
// gather Y into the Hermitian full spectrum fr [B][L]
void launch_cqt_synth_gather(const CqtTables& t, int B, const float2* Y, float2* fr, cudaStream_t s);
void launch_spec_mul_real(int B, int L, float2* spec, const float* h, cudaStream_t s);
// adjoint pieces (backward.cu): full-circle gather without Hermitian completion (t.dual = analysis windows / M); spectrum factors of
// the synthesis adjoint
void launch_cqt_gather_adj(const CqtTables& t, int B, int oct_lo, int oct_hi, const float2* Y, float2* X, bool accumulate, cudaStream_t s);
void launch_spec_synth_adj(int B, int L, float2* G, float scale, cudaStream_t s);

// ---- backward (input gradient) kernels, backward.cu ---------------------------------------------------
void launch_scale_channels(const TV& in, const float* vec, long long vstride, float alpha, const TV& out, cudaStream_t s);
// out = cres * gres + d/dx [act(GroupNorm(x) * (affine + 1))]^T g ; D_scratch: [B][8] doubles; gres.p may be null; out may alias gres / g
void launch_gn_bwd(const TV& g, const TV& x, const double* stats, long long n_per_group, const float* gamma, const float* affine,
                   long long abstride, bool gelu, double* D_scratch, const TV& gres, float cres, const TV& out, cudaStream_t s,
                   unsigned int* amax_out = nullptr);
// per-tensor fp16 scaling of a gradient operand: amax accumulates max |x| (float bits); tc_scale turns it into the operand
// multiplier scal[0] = kappa * pre and the epilogue vector inv_vec[0..n) = post / kappa, kappa a power of two, and clears amax
void launch_absmax(const TV& x, unsigned int* amax, cudaStream_t s);
void launch_tc_scale(unsigned int* amax, float pre, float post, float* scal, float* inv_vec, int n, cudaStream_t s);
void launch_transpose_weight_std(const float* wp, float* wstd, int Cout, int Cin, int taps, cudaStream_t s);
// gx = beta * gx + adjoint of the resampler applied to gy (gx: the resampler's input shape, gy: its output shape)
void launch_resample_down_adj(const TV& gy, const TV& gx, float beta, cudaStream_t s);
void launch_resample_up_adj(const TV& gy, const TV& gx, float beta, cudaStream_t s);
struct BGemm {   // C[z](m,n) = alpha * sum_k A[z](m,k) B[z](k,n) + beta * C[z](m,n); element strides
    const float* A; const float* B; float* C;
    int M, N, K, batch;
    long long sAm, sAk, sAz, sBk, sBn, sBz, sCm, sCn, sCz;
    float alpha, beta;
};
void launch_bgemm(const BGemm& p, cudaStream_t s);
void launch_softmax_rows(float* S, long long rows, int T, cudaStream_t s);
void launch_softmax_bwd(const float* P, float* gP, long long rows, int T, float scale, cudaStream_t s);
void launch_pack_conv_weight_T(const float* wp, float* wpT, int Cout, int Cin, int KF, int KT, cudaStream_t s);

// EDM sampler element-wise steps (sampler.py:214, 141-147, 230-251)
void launch_axpy_noise(float* x, const float* eps, float scale, long long n, cudaStream_t s);
// sigma_h_dev != null: sigma = sigma_h_dev[0], h = sigma_h_dev[1] (device scalars, CUDA-graph replay)
void launch_edm_step(const float* xin, const float* xhat, const float* y, const float* mask, long long mask_n, long long n,
                     float sigma, float h, int mode, const float* d_prev, const float* xbase, float* d_out, float* x_out,
                     cudaStream_t s, const float* sigma_h_dev = nullptr);
// Philox4x32-10 standard normals keyed by (seed, stream, global clip index, draw): x[c][i] = (accumulate ? x : 0) + scale * N(0,1).
// scale_draw_dev != null: [scale, draw, stream_id, clip0] (the last three as bit patterns) come from device memory instead
void launch_philox_normal(float* x, int n_clips, long long L, unsigned long long seed, unsigned int stream_id, unsigned int clip0,
                          unsigned int draw, float scale, bool accumulate, const float* scale_draw_dev, cudaStream_t s);
// cur[0..row) = table[*counter][0..row); ++*counter   (one tiny launch at the head of a replayed sampler-step graph)
void launch_sched_select(const float* table, int row, int* counter, float* cur, cudaStream_t s);

}  // namespace aid
