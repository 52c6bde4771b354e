// fp32 CUDA-core implicit-GEMM convolution with a fused residual/gate/statistics epilogue (sm_100a).
//
// Replaces F.conv2d(x, w, padding="same", dilation=(d,1)) of unet.py:85 for the 5x3 dilated layers
// (unet.py:433-436, 482), every 1x1 projection (unet.py:322-323, 412-415), pyr_down_proj (unet.py:676, 794)
// and the attention qk Conv1d (unet.py:321, 355; viewed as a 1x1 conv over [B, 8F, 1, T]).
//
// GEMM view: M = pixels of one clip (tile of 128, PT_T along T x 128/PT_T along F), N = Cout (tile 64),
// K = Cin * KF * KT, gathered on the fly into shared memory (zero "same" padding = bounds check).
// Weights are pre-packed K-major: wp[(ci*KF*KT + tap) * Cout + co].
// This is the exact-fp32 path: it is the parity reference for the tensor-core path and runs every
// convolution that the tcgen05 kernel does not cover.
#include "common.cuh"

namespace aid {

static constexpr int BM = 128;  // pixels per CTA
static constexpr int BN = 64;   // output channels per CTA
static constexpr int NT = 256;  // threads: 16 pixel groups (8 px) x 16 channel groups (4 co)

template <int KF, int KT, int CI_CHUNK>
__global__ void __launch_bounds__(NT)
conv_simt_kernel(TV a, const float* __restrict__ wp, int dil, TV out, ConvEpilogue ep, int pt_t_log2, int tiles_t) {
    constexpr int TAPS = KF * KT;
    constexpr int KC = CI_CHUNK * TAPS;
    extern __shared__ float smem[];
    float* As = smem;             // [KC][BM]
    float* Ws = smem + KC * BM;   // [KC][BN]
    __shared__ double sstat[8][2];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int co0 = blockIdx.y * BN;
    const int pt_t = 1 << pt_t_log2, pt_f = BM >> pt_t_log2;
    const int tile_t = blockIdx.x % tiles_t, tile_f = blockIdx.x / tiles_t;
    const int f0 = tile_f * pt_f, t0 = tile_t * pt_t;
    const int Cin = a.C, Cout = out.C, Fd = a.F, T = a.T;
    const int Ktot = Cin * TAPS;

    // gather role: fixed pixel p = tid % 128, k index advances by 2 per iteration
    const int gp = tid & (BM - 1);
    const int gf = f0 + (gp >> pt_t_log2), gt = t0 + (gp & (pt_t - 1));
    const float* abase = a.p + (long long)b * a.sb;

    // compute role
    const int py = tid >> 4, cx = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int ci0 = 0; ci0 < Cin; ci0 += CI_CHUNK) {
        // ---- gather A chunk: As[kk][p] = a[b, ci0 + kk/TAPS, f + (kf-KF/2)*dil, t + kt - KT/2] ----
#pragma unroll 4
        for (int kk = tid >> 7; kk < KC; kk += NT / BM) {
            const int ci = ci0 + kk / TAPS, tap = kk % TAPS;
            const int kf = tap / KT, kt = tap % KT;
            const int ff = gf + (kf - KF / 2) * dil, tt = gt + kt - KT / 2;
            float v = 0.f;
            if (ci < Cin && ff >= 0 && ff < Fd && tt >= 0 && tt < T)
                v = __ldg(abase + (long long)ci * a.sc + (long long)ff * T + tt);
            As[kk * BM + gp] = v;
        }
        // ---- load W chunk: Ws[kk][co] = wp[(ci0*TAPS + kk) * Cout + co0 + co] ----
        for (int e = tid; e < KC * BN; e += NT) {
            const int kk = e >> 6, co = e & 63;
            const int k = ci0 * TAPS + kk;
            float v = 0.f;
            if (k < Ktot && co0 + co < Cout) v = __ldg(wp + (long long)k * Cout + co0 + co);
            Ws[e] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + kk * BM + py * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(As + kk * BM + py * 8 + 4);
            const float4 w = *reinterpret_cast<const float4*>(Ws + kk * BN + cx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: out = alpha*(acc*gate + R) + beta*R2, statistics of out ----
    if (ep.stats && tid < 16) sstat[tid >> 1][tid & 1] = 0.0;
    if (ep.stats) __syncthreads();
    const int p0 = py * 8;  // 8 consecutive pixels, same row because pt_t >= 8
    const int f = f0 + (p0 >> pt_t_log2), tb = t0 + (p0 & (pt_t - 1));
    const int gcn = Cout >= 8 ? Cout / 8 : 1;
    float ssum = 0.f, ssq = 0.f;
    int sgroup = -1;
    const bool per_channel_stats = ep.stats && (gcn & 3) != 0;  // a thread's 4 channels may straddle groups
    if (f < Fd) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + cx * 4 + j;
            if (co >= Cout) continue;
            if (per_channel_stats && sgroup >= 0 && sgroup != co / gcn) {
                atomicAdd(&sstat[sgroup & 7][0], (double)ssum);
                atomicAdd(&sstat[sgroup & 7][1], (double)ssq);
                ssum = 0.f; ssq = 0.f;
            }
            const float g = ep.gate ? ep.gate[(long long)b * ep.gate_bstride + co] : 1.f;
            const long long po = (long long)b * out.sb + (long long)co * out.sc + (long long)f * T;
            const long long pr = ep.R.p ? (long long)b * ep.R.sb + (long long)co * ep.R.sc + (long long)f * T : 0;
            const long long pr2 = ep.R2.p ? (long long)b * ep.R2.sb + (long long)co * ep.R2.sc + (long long)f * T : 0;
            sgroup = co / gcn;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = tb + i;
                if (t >= T) continue;
                float v = acc[i][j] * g;
                if (ep.R.p) v += ep.R.p[pr + t];
                v *= ep.alpha;
                if (ep.R2.p) v += ep.beta * ep.R2.p[pr2 + t];
                out.p[po + t] = v;
                ssum += v; ssq += v * v;
            }
        }
    }
    if (ep.stats) {
        // the 4 channels of a thread share a group (host guarantees (Cout/8) % 4 == 0 when stats are requested)
        if (sgroup >= 0) {
            atomicAdd(&sstat[sgroup & 7][0], (double)ssum);
            atomicAdd(&sstat[sgroup & 7][1], (double)ssq);
        }
        __syncthreads();
        // groups touched by this CTA: co0/gcn .. (co0+BN-1)/gcn  (at most 8 distinct, indexed modulo 8)
        if (tid < 8) {
            const int glo = co0 / gcn, ghi = min(7, (min(Cout, co0 + BN) - 1) / gcn);
            for (int g = glo; g <= ghi; ++g)
                if ((g & 7) == tid) {
                    atomicAdd(ep.stats + ((long long)b * 8 + g) * 2 + 0, sstat[tid][0]);
                    atomicAdd(ep.stats + ((long long)b * 8 + g) * 2 + 1, sstat[tid][1]);
                }
        }
    }
}

template <int KF, int KT, int CI_CHUNK>
static void launch_t(const TV& a, const float* wp, int dil, const TV& out, const ConvEpilogue& ep, cudaStream_t s) {
    constexpr int KC = CI_CHUNK * KF * KT;
    const size_t smem = (size_t)KC * (BM + BN) * sizeof(float);
    static SmemConfig configured;
    ensure_dyn_smem(conv_simt_kernel<KF, KT, CI_CHUNK>, smem, configured);
    int pt_t = 8, lg = 3;
    while (pt_t < BM && pt_t < a.T) { pt_t <<= 1; ++lg; }
    const int pt_f = BM / pt_t;
    const int tiles_t = (a.T + pt_t - 1) / pt_t, tiles_f = (a.F + pt_f - 1) / pt_f;
    dim3 grid(tiles_t * tiles_f, (out.C + BN - 1) / BN, a.B);
    conv_simt_kernel<KF, KT, CI_CHUNK><<<grid, NT, smem, s>>>(a, wp, dil, out, ep, lg, tiles_t);
    AID_COUNT_LAUNCH(1);
}

void launch_conv_simt(const TV& a, const float* wp, int KF, int KT, int dil, const TV& out, const ConvEpilogue& ep,
                      cudaStream_t s) {
    if (KF == 5 && KT == 3) launch_t<5, 3, 8>(a, wp, dil, out, ep, s);
    else if (KF == 1 && KT == 1) launch_t<1, 1, 32>(a, wp, 1, out, ep, s);
    else throw CudaError(cudaErrorInvalidValue, "unsupported conv kernel size", __FILE__, __LINE__);
}

}  // namespace aid
