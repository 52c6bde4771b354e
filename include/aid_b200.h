/* C ABI of the B200-native denoiser path (libaid_b200.so).
 *
 * Drop-in boundary for ONE path of eloimoliner/audio-inpainting-diffusion: the denoiser forward
 * driven by the EDM inpainting sampler.  The reference is pure Python; the "FFI" a maintainer would
 * bind is ctypes.  Every entry point names the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only.  `*_dev` pointers are device pointers on the GPU the
 * handle was created for; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * All entry points return 0 on success or a negative aid_status; aid_last_error() gives the message.
 * Calls that take a handle are stream-ordered, never synchronise the device and never allocate device memory after
 * aid_finalize() (the caller owns the workspace); the one exception is the CQT entry points used BEFORE aid_finalize,
 * which upload their tables on first use.  The handle-less single-operator and aid_debug_* entry points exist for unit
 * parity tests and tuning: they allocate and free their own scratch and may synchronise.  The current CUDA device of the
 * calling thread is left as it was found.  One handle = one GPU = one host thread at a time.
 */
#ifndef AID_B200_H
#define AID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AID_MAX_OCTS 16

typedef enum aid_status {
    AID_OK = 0,
    AID_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
    AID_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    AID_ERR_STATE = -3,     /* wrong call order (e.g. denoise before finalize, missing weights) */
    AID_ERR_WORKSPACE = -4  /* caller-provided workspace too small */
} aid_status;

/* Mirrors what Unet_CQT_oct_with_attention.__init__ reads from `args` (unet.py:595-655):
 * args.network.{cqt.num_octs,cqt.bins_per_oct,cqt.window,cqt.beta,emb_dim,Ns,num_dils,attention_layers,
 * attention_dict.num_heads,num_bottleneck_layers} and args.exp.{sample_rate,audio_len}. */
typedef struct aid_config {
    int32_t num_octs;
    int32_t bins_per_oct;
    int32_t audio_len;            /* power of two >= 4096 */
    int32_t window_kind;          /* 0 = "hann", 1 = ("kaiser", beta) */
    double sample_rate;
    double beta;
    int32_t emb_dim;              /* must be 256 (RFF_MLP_Block 64->128->256->emb_dim) */
    int32_t num_heads;            /* attention_dict.num_heads */
    int32_t Ns[AID_MAX_OCTS];
    int32_t num_dils[AID_MAX_OCTS];
    int32_t attention_layers[AID_MAX_OCTS + 1]; /* num_octs entries + bottleneck */
    int32_t num_bottleneck_layers; /* must be 1 */
    int32_t conv_mode;            /* 0 = exact fp32 CUDA cores; 1 = tcgen05, split-fp16 operands (fp32-grade); 2 = tcgen05, single fp16 operands */
} aid_config;

typedef struct aid_handle aid_handle;

/* Unet_CQT_oct_with_attention(args, device)                                   unet.py:587 */
int aid_create(const aid_config* cfg, int device, aid_handle** out);
void aid_destroy(aid_handle* h);
const char* aid_last_error(const aid_handle* h); /* h may be NULL: error of the last failed aid_create */

/* model.load_state_dict(sd): one call per tensor, reference key names   tester_inpainting.py:202
 * `host` is a host pointer to contiguous fp32 data. */
int aid_load_weight(aid_handle* h, const char* name, const float* host, const int64_t* shape, int ndim);
/* number of tensors the schema expects / names by index (for state_dict()/strict loading) */
int aid_num_weights(const aid_handle* h);
int aid_weight_info(const aid_handle* h, int index, const char** name, int64_t* shape4, int* ndim);
/* repack weights into kernel layouts and upload; required before any compute call */
int aid_finalize(aid_handle* h);

/* bytes of caller-owned scratch aid_unet_forward needs for a batch of B clips */
int aid_workspace_bytes(aid_handle* h, int B, size_t* bytes);

/* Unet_CQT_oct_with_attention.forward(inputs[B,L], sigma[B|1,1]) -> [B,L]      unet.py:730-845
 * fused with EDM.denoiser's preconditioning (edm.py:133-148):
 *     out = out_scale * net(in_scale * x, c_noise) + skip_scale * x
 * (in_scale=1, out_scale=1, skip_scale=0 is the bare nn.Module forward).
 * c_noise_dev holds n_sigma (1 or B) values.  out_dev may alias x_dev only if skip_scale == 0. */
int aid_unet_forward(aid_handle* h, const float* x_dev, const float* c_noise_dev, int n_sigma, float* out_dev, int B,
                     float in_scale, float out_scale, float skip_scale, void* workspace_dev, size_t workspace_bytes,
                     void* stream);

/* The same forward with the preconditioning scalars in DEVICE memory: scales_dev = [in_scale, out_scale, skip_scale].  Nothing
 * of the current noise level is baked into the launch parameters, so a captured CUDA graph of a sampler step can be replayed for
 * every step of the schedule (sampler.py:201-251); out_dev may not alias x_dev. */
int aid_unet_forward_ds(aid_handle* h, const float* x_dev, const float* c_noise_dev, int n_sigma, float* out_dev, int B,
                        const float* scales_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Input gradient of the denoiser (reconstruction guidance, sampler.py:57-113: torch.autograd.grad(norm, x) through
 * EDM.denoiser, edm.py:133-148, and Unet_CQT_oct_with_attention.forward, unet.py:730-845).  Weights get no gradient.
 *   aid_unet_forward_tape : the forward of aid_unet_forward, keeping in the workspace everything the backward needs (every block
 *                           and layer input, group statistics, attention q|k); workspace size from aid_vjp_workspace_bytes.
 *   aid_unet_backward     : grad_x = in_scale * out_scale * J_net(in_scale * x)^T grad_out + skip_scale * grad_out for the taped
 *                           forward.  The workspace must be untouched since aid_unet_forward_tape; may be called repeatedly.
 * The first call on a handle builds the transposed convolution weights and the CQT adjoint tables (one device allocation each).
 * Backward arithmetic is fp32 on CUDA cores in every conv_mode; the taped forward runs in the handle's conv_mode. */
int aid_vjp_workspace_bytes(aid_handle* h, int B, size_t* bytes);
int aid_unet_forward_tape(aid_handle* h, const float* x_dev, const float* c_noise_dev, int n_sigma, float* out_dev, int B,
                          float in_scale, float out_scale, float skip_scale, void* workspace_dev, size_t workspace_bytes,
                          void* stream);
int aid_unet_backward(aid_handle* h, const float* grad_out_dev, float* grad_x_dev, void* stream);

/* CQT_nsgt.fwd / bwd / apply_hpf_DC                         unet.py:743, unet.py:841, sampler.py:63,123
 * coefficient layout: per octave o (ascending frequency) a [B, 2, bins, T_o] fp32 tensor (re, im planes),
 * octave tensors concatenated in one buffer at offsets aid_cqt_layout() reports (in floats, for batch B). */
int aid_cqt_layout(const aid_handle* h, int B, int64_t* offsets /*[num_octs+1]*/, int32_t* frames /*[num_octs]*/);
int aid_cqt_fwd(aid_handle* h, const float* x_dev, float* coef_dev, int B, void* workspace_dev, size_t workspace_bytes, void* stream);
int aid_cqt_bwd(aid_handle* h, const float* coef_dev, float* x_dev, int B, void* workspace_dev, size_t workspace_bytes, void* stream);
int aid_hpf_dc(aid_handle* h, const float* x_dev, float* out_dev, int B, void* workspace_dev, size_t workspace_bytes, void* stream);
int aid_cqt_workspace_bytes(const aid_handle* h, int B, size_t* bytes);
/* host-side band plan (for inspection / tests): K bands, n_win window samples in total.
 * Any output pointer may be NULL.  centre/Lg/woff: [K]; win/dual: [n_win]; hhpf: [audio_len]. */
int aid_cqt_plan(const aid_handle* h, int32_t* K, int32_t* n_win, int32_t* centre, int32_t* Lg, int32_t* woff,
                 float* win, float* dual, float* hhpf);

/* Sampler element-wise steps on [B, L] buffers                          sampler.py:214, 141-147, 230-251 */
int aid_edm_add_noise(float* x_dev, const float* eps_dev, float scale, int64_t n, void* stream);
/* xh = mask ? mask*y + (1-mask)*xhat : xhat;  d = (xin - xh)/sigma;
 * mode 0: d_out = d (if non-NULL), x_out = xin + h*d;   mode 1: x_out = xbase + h*(d_prev + d)/2 */
int aid_edm_step(const float* xin_dev, const float* xhat_dev, const float* y_dev, const float* mask_dev, int64_t mask_n,
                 int64_t n, float sigma, float h, int mode, const float* d_prev_dev, const float* xbase_dev,
                 float* d_out_dev, float* x_out_dev, void* stream);

/* aid_edm_step with sigma = sigma_h_dev[0], h = sigma_h_dev[1] read on the device (graph replay) */
int aid_edm_step_ds(const float* xin_dev, const float* xhat_dev, const float* y_dev, const float* mask_dev, int64_t mask_n, int64_t n,
                    const float* sigma_h_dev, int mode, const float* d_prev_dev, const float* xbase_dev, float* d_out_dev,
                    float* x_out_dev, void* stream);

/* Device-resident noise for sample_prior (edm.py:87-95) and the churn noise (sampler.py:210-214), which the reference draws on the
 * CPU generator and copies (4*B*L bytes per step): x[c][i] = (accumulate ? x[c][i] : 0) + scale * n(seed, stream_id, clip0 + c, draw, i)
 * for c < n_clips, i < L, where n() are standard normals from Philox4x32-10 (counter = (i/4, draw, clip, stream_id), key = seed) and
 * Box-Muller.  A clip's noise does not depend on the batch or rank it is sampled in.  scale_draw_dev (may be NULL): device scalars
 * [scale, draw, stream_id, clip0] (the last three as 32-bit patterns) overriding the by-value arguments (graph replay).  Restated in oracle/philox_oracle.py. */
int aid_philox_normal(float* x_dev, int n_clips, int64_t L, uint64_t seed, uint32_t stream_id, uint32_t clip0, uint32_t draw,
                      float scale, int accumulate, const float* scale_draw_dev, void* stream);

/* cur_dev[0..row_floats) = table_dev[*counter_dev][0..row_floats); ++*counter_dev -- the head node of a replayed sampler-step graph:
 * row k of the host-built schedule table (noise scale, draw, c_in / c_out / c_skip / c_noise / sigma / h of the step's evaluations)
 * becomes the current device scalars.  row_floats <= 64. */
int aid_sched_select(const float* table_dev, int row_floats, int32_t* counter_dev, float* cur_dev, void* stream);

/* Spectrogram-inpainting degradation S(x) = crop(istft(mask * stft(zero-pad(x)))) with a periodic Hann window of n_fft samples,
 * centre = True / reflect padding, as Sampler.apply_spectral_mask does with torch.stft / torch.istft   sampler.py:271-290
 * y_dev == NULL: out = S(x);  else out = y + x - S(x), the projection of the spectrogram mode          sampler.py:361
 * mask_dev: float [n_fft/2 + 1][n_frames], n_frames = 1 + (L + n_fft - L % n_fft) / hop;  n_fft a power of two <= 4096;
 * frames_dev: scratch of B * n_frames * n_fft floats.  x, y, out: [B, L]; out may alias neither x nor y. */
int aid_spectral_mask(const float* x_dev, const float* y_dev, const float* mask_dev, int B, int64_t L, int n_fft, int hop,
                      int n_frames, float* frames_dev, size_t frames_bytes, float* out_dev, void* stream);

/* ---- single-operator entry points (unit parity tests; same kernels the forward uses) ---------------- */
/* F.conv2d(a[B,Cin,F,T], w[Cout,Cin,KF,KT], padding="same", dilation=(dil,1)) with the fused epilogue
 * out = alpha*(conv*gate[c] + R) + beta*R2; gate/R/R2 may be NULL.  stats_dev (may be NULL): [B][8][2] doubles
 * accumulated with (sum, sumsq) of out per channel group.  mode: 0 = fp32 CUDA cores (thin-channel kernels where they apply),
 * 1 = tcgen05 split-fp16 (3 MMAs per tap), 2 = force the general fp32 CUDA-core kernel, 3 = tcgen05 single fp16 (1 MMA per tap),
 * 4 = like 3 with a, R and out given channels-last ([B,F,T,C], the layout the residual stream has inside a conv_mode 2 block).
 *                                                                                            unet.py:79-88, 482 */
int aid_op_conv2d(const float* a_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                  const float* gate_dev, const float* R_dev, const float* R2_dev, float alpha, float beta,
                  float* out_dev, double* stats_dev, int mode, void* stream);
/* BiasFreeGroupNorm (8 groups) * (affine+1) [+ exact GELU]                       unet.py:147-163, 479, 482 */
int aid_op_groupnorm_act(const float* x_dev, const float* gamma_dev, const float* affine_dev, int B, int C, int F, int T,
                         int gelu, float* out_dev, double* stats_scratch_dev /*[B][8][2]*/, void* stream);
/* UpDownResample along T: up=0 -> [.., T/2], up=1 -> [.., 2T]                               unet.py:549-580 */
int aid_op_resample(const float* x_dev, int B, int C, int F, int T, int up, float* out_dev, void* stream);
/* attention core of TimeAttentionBlock: h[B,heads,F,T], qk[B,2*heads*F,T] -> out[B,heads,F,T]   unet.py:353-374 */
int aid_op_attention(const float* h_dev, const float* qk_dev, int B, int heads, int F, int T, float* out_dev, void* stream);
/* the same with the kernel chosen: tensor_core = 0 fp32 CUDA cores (conv_mode 0 / 1), 1 = tcgen05 (conv_mode 2: Q K^T with split-fp16
 * operands accumulated in TMEM, softmax read from TMEM, P V with fp16 operands; T <= 256, F % 16 == 0, F <= 512) */
int aid_op_attention_mode(const float* h_dev, const float* qk_dev, int B, int heads, int F, int T, float* out_dev, int tensor_core,
                          void* stream);
/* RFF_MLP_Block + every adaLN Linear: returns the emb [n_sigma,256]                          unet.py:184-211 */
int aid_op_embedding(aid_handle* h, const float* c_noise_dev, int n_sigma, float* emb_dev, void* stream);

/* Backward (input-gradient) single operators, the pieces of aid_unet_backward:
 *   aid_op_resample_adj       adjoint of aid_op_resample: gy [.., up ? 2T : T/2] -> gx [.., T]                    unet.py:549-580
 *   aid_op_groupnorm_act_bwd  gx = d/dx [act(GroupNorm(x) * (affine + 1))]^T g; scratch_dev: [B][8][3] doubles     unet.py:147-163
 *   aid_op_attention_bwd      g_o [B,heads,F,T] -> gh (value path only) [B,heads,F,T] and gqk [B,2*heads*F,T];
 *                             scratch: 2 * B * heads * T * T floats (+ alignment)                               unet.py:353-374
 *   aid_op_conv2d_bwd_input   gx = conv^T(gy): the forward kernels with tap-mirrored, channel-swapped weights      unet.py:79-88
 *   aid_cqt_fwd_vjp / aid_cqt_bwd_vjp  adjoints of aid_cqt_fwd / aid_cqt_bwd (same layout and workspace; after aid_finalize) */
int aid_op_resample_adj(const float* gy_dev, int B, int C, int F, int T, int up, float* gx_dev, void* stream);
int aid_op_groupnorm_act_bwd(const float* g_dev, const float* x_dev, const float* gamma_dev, const float* affine_dev, int B, int C, int F,
                             int T, int gelu, float* gx_dev, double* scratch_dev, void* stream);
int aid_op_attention_bwd(const float* h_dev, const float* qk_dev, const float* go_dev, int B, int heads, int F, int T, float* gh_dev,
                         float* gqk_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
int aid_op_conv2d_bwd_input(const float* gy_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                            float* gx_dev, void* stream);
int aid_cqt_fwd_vjp(aid_handle* h, const float* gcoef_dev, float* gx_dev, int B, void* workspace_dev, size_t workspace_bytes, void* stream);
int aid_cqt_bwd_vjp(aid_handle* h, const float* gx_dev, float* gcoef_dev, int B, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Debug/tuning (not part of the drop-in surface): aid_op_conv2d run twice, returning the device time in ms of the
 * second convolution launch alone. */
int aid_debug_time_conv2d(const float* a_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                          const float* gate_dev, const float* R_dev, float alpha, float* out_dev, double* stats_dev, int mode,
                          float* ms_out);

/* Debug / layout parity (not part of the drop-in surface): writes the conv_mode 2 tensor-core operand layouts of an
 * activation x[B,Cin,F,T] (channels-last fp16, swizzled, PF pad rows) and of a weight w[Cout,Cin,KF,KT]; sizes in halves. */
int aid_debug_tc2_operands(const float* x_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int PF,
                           void* a_out_dev, void* w_out_dev, uint64_t* a_halves, uint64_t* w_halves);

/* Debug / tuning: with AID_TC_DEBUG bit 2048 set, conv_tc2_kernel sums the cycles its warp roles spend waiting and working:
 * [0] MMA wait tmem_empty, [1] MMA wait a_full, [2] MMA wait b_full, [3] MMA issue + commit, [4] MMA total, [5] A producer wait
 * a_empty, [6] B producer wait b_empty, [7] epilogue warp 0 wait tmem_full, [8] epilogue warp 0 total, [9] CTAs.  Read and cleared. */
int aid_debug_tc2_profile(uint64_t* out16);

/* Debug / tuning: ms per launch of the conv_mode 2 normalise + GELU + operand-layout pass on x[B,C,F,T]. */
int aid_debug_time_gn_tc2(const float* x_dev, int B, int C, int F, int T, int PF, int iters, float* ms_out);

/* Debug / parity / tuning of one dilated residual layer of conv_mode 2 (unet.py:470-482),
 *   out = alpha * (x + gate[c] * conv5x3_dil(GELU(GroupNorm8(x) * gamma * (1 + affine))))          x, out: [B, C, F, T], w: [C, C, 5, 3]
 * fused = 0: operand pass (gn_act_tc2) + conv_tc2_kernel;  fused = 1: conv_comb_kernel (normalisation, GELU and operand conversion
 * inside the convolution; C = 64 or 96, T % 128 == 0).  stats_out_dev ([B][8][2] doubles, may be NULL) receives (sum, sumsq) of out per
 * group.  ms_out (may be NULL): device time of the layer (pass + convolution, or the fused kernel) of a second, timed run. */
int aid_debug_dilated_layer(const float* x_dev, const float* w_dev, int B, int C, int F, int T, int dil, const float* gamma_dev,
                            const float* affine_dev, const float* gate_dev, float alpha, int fused, float* out_dev, double* stats_out_dev,
                            float* ms_out);

/* debug / parity / tuning: the init block of an encoder level in conv_mode 2 (reference unet.py:452-493 as instantiated at unet.py:673-675:
 * dim = 2 CQT channels, one gated 1x1 residual layer, res_conv):  y = proj_in(x2);  x = (y + gate * H(GELU(GN8(y) gamma (1 + affine)))) / sqrt 2;
 * out = (x + res_conv(x2)) / sqrt 2.  fused = 0: the five launches of the un-fused path;  fused = 1: init_block_kernel (conv_init.cu; N = 64, 96
 * or 128, T % 128 == 0), which reads x2 once and never materialises y.  Weights in the checkpoint layout (w_in, w_res: [N][2], wH: [N][N]);
 * gamma, affine, gate: [N].  stats_out_dev ([B][8][2] doubles, may be NULL): (sum, sumsq) of out per group.  ms_out (may be NULL): device time
 * of a second, timed run. */
int aid_debug_init_block(const float* x2_dev, const float* w_in_dev, const float* w_res_dev, const float* wH_dev, int B, int N, int F, int T,
                         const float* gamma_dev, const float* affine_dev, const float* gate_dev, int fused, float* out_dev,
                         double* stats_out_dev, float* ms_out);

/* debug / parity / tuning: an out block of the decoder / bottleneck in conv_mode 2 (reference unet.py:452-493 as instantiated at unet.py:690, 719:
 * one gated 1x1 residual layer N -> N, then proj_out and res_conv N -> 2):  x1 = (x + gate * H(GELU(GN8(x) gamma (1 + affine)))) / sqrt 2;
 * blk = (proj_out(x1) + res_conv(x)) / sqrt 2;  out = blk, or (accum + blk) / sqrt 2 when accum_dev != NULL (unet.py:817; out_dev may alias it).
 * fused = 0: the four launches of the un-fused path (fp16 operands for H);  fused = 1: out_block_kernel (out_block.cu), which folds
 * P diag(gate) H into a per-clip 2 x N matrix and reads x once, all in fp32.  Weights in the checkpoint layout (wH: [N][N], wP, wR: [2][N]);
 * gamma, affine, gate: [N].  ms_out (may be NULL, needs accum_dev == NULL): device time of a second, timed run. */
int aid_debug_out_block(const float* x_dev, const float* wH_dev, const float* wP_dev, const float* wR_dev, int B, int N, int F, int T,
                        const float* gamma_dev, const float* affine_dev, const float* gate_dev, const float* accum_dev, int fused, float* out_dev,
                        float* ms_out);

/* debug / parity: choose between the fused kernels of conv_mode 2 and their un-fused twins for later forwards of this handle
 * (1 = fused, the default; 0 = un-fused; -1 = leave).  init_blocks: init_block_kernel vs five launches; dilated_layers: conv_comb_kernel /
 * conv_comb96_kernel vs operand pass + conv_tc2_kernel; out_blocks: out_block_kernel vs four launches; upsampling: the decoder stream upsampled
 * inside the operand conversion of the next level's main block vs resample_up + conversion of its fp32 copy. */
int aid_debug_fusion(aid_handle* h, int init_blocks, int dilated_layers, int out_blocks, int upsampling);

/* Per-launch timing of the convolution kernels with CUDA events on the launching stream (bench.py's roofline).
 * aid_profile(h, 1) clears and starts recording, aid_profile(h, 0) stops; aid_profile_read sums the recorded launches
 * of one kind (0 = dilated 5x3 residual-layer convolutions on conv_tc2 / conv_tc / conv_simt, 1 = all other convolutions, 2 = fused
 * dilated residual layers of conv_comb.cu, whose time includes normalisation, GELU and operand conversion): count, device ms,
 * executed FLOPs (2*Cin*Cout*taps*pixels, row taps outside the plane not counted) and algorithmic bytes (kind 0 / 1: operand +
 * result + residual tensors + weights in fp32; kind 2: 4 B read + 4 B written per element + fp16 weights). */
int aid_profile(aid_handle* h, int enable);
int aid_profile_read(aid_handle* h, int kind, uint64_t* launches, double* ms, double* flops, double* bytes);

/* Debug: the next aid_unet_forward calls also copy the named intermediate, contiguous [B,C,F,T], into dst_dev
 * (NULL removes the probe).  Names: "enc<i>" = encoder ResBlock output of level i (unet.py:780), "mid" = bottleneck
 * ResBlock output (unet.py:803), "dec<i>" = decoder ResBlock output (unet.py:815). */
int aid_debug_probe(aid_handle* h, const char* name, float* dst_dev);

/* Debug: conv_mode 2 converts its operands to fp16 with saturation (activations x 2^4, weights x 2^10, finite range 65504).
 * enable != 0 resets the activation counter and counts, in the following forwards, every operand value the
 * normalise/GELU/convert passes had to clamp; a non-NULL act_count reads it (synchronises the device).  weight_count:
 * weight values clamped when aid_finalize packed them.  Both are 0 in conv_mode 0 / 1 (nothing is clamped there). */
int aid_debug_saturation(aid_handle* h, int enable, uint64_t* act_count, uint64_t* weight_count);

/* kernels launched by this library since load (the bench's gpu_launches counter) */
uint64_t aid_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AID_B200_H */
