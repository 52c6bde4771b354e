// Fused time-attention core (unet.py:353-374): per (clip, head)
//   q[t, d] = qk[b, head*2F + d, t],  k[t, d] = qk[b, head*2F + F + d, t],  v[t, f] = h[b, head, f, t]
//   out[b, head, f, t] = sum_tk softmax_tk(F^-0.5 * q[t,:].k[tk,:]) * v[tk, f]
// One CTA handles TQ queries of one (clip, head): scores live in shared memory (T <= ~1000 frames).
// fp32 CUDA-core version (exact-parity path; < 0.1 % of the forward's FLOPs).
#include "common.cuh"

namespace aid {

static constexpr int TQ = 16;     // queries per CTA
static constexpr int FCH = 32;    // value rows staged per pass
static constexpr int ANT = 256;

// Shared memory: qs [F][TQ] (scores phase) is reused as the P.V workspace afterwards.
//   S   [TQ][T+1]   scores / probabilities, row per query (conflict-free for the score and softmax sweeps)
//   PT  [T][TQ]     probabilities transposed: one 64-byte row per key
//   hT  [T][FCH]    value rows transposed, 16-byte groups XOR-swizzled by (key & 7) so that the transposing stores spread
//                   over the banks while the P.V loop still reads aligned float4s
//   red [8][FCH*TQ] per-warp partial outputs (each warp sums over an eighth of the keys)
// P.V register tile: a thread owns 4 value rows x 4 queries; per key it reads one float4 of hT and one float4 of PT for 16
// FMAs (the first version read 3 scalars for 2 FMAs and spent ~90 % of its time in the shared-memory pipe).
__global__ void __launch_bounds__(ANT)
attention_kernel(TV h, const float* __restrict__ qk, TV out, float scale) {
    extern __shared__ __align__(16) float sm[];
    const int F = h.F, T = h.T;
    const int Tp = T + 1;
    float* qs = sm;                                   // [F][TQ]
    float* S = qs + (size_t)F * TQ;                   // [TQ][Tp]
    float* PT = S + (((size_t)TQ * Tp + 3) & ~(size_t)3);   // [T][TQ]
    float* hT = PT + (size_t)T * TQ;                  // [T][FCH]
    float* red = hT + (size_t)T * FCH;                // [8][FCH * TQ]; during the scores phase: partial scores of the d-splits 1..3

    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int head = blockIdx.y, b = blockIdx.z;
    const int t0 = blockIdx.x * TQ;
    const float* qbase = qk + ((long long)b * h.C * 2 * F + (long long)head * 2 * F) * T;
    const float* kbase = qbase + (long long)F * T;

    // stage q tile
    for (int e = tid; e < F * TQ; e += ANT) {
        const int d = e / TQ, j = e % TQ;
        qs[e] = (t0 + j < T) ? __ldg(qbase + (long long)d * T + t0 + j) : 0.f;
    }
    __syncthreads();

    // scores: thread <-> (key column, slice of the feature dimension).  With fewer keys than threads (T = 64, 128 on the
    // deepest levels) the threads of one key split d between them; the partial scores of splits 1.. go to `red` (free until the
    // P.V phase) and are added in a fixed order, so the result is reproducible run to run (no floating-point atomics).
    const int nsplit = (T < ANT && ANT % T == 0 && (ANT / T - 1) * TQ * Tp <= (ANT / 32) * FCH * TQ) ? ANT / T : 1;
    for (int tk = tid % (nsplit > 1 ? T : ANT); tk < T; tk += ANT) {
        const int part = nsplit > 1 ? tid / T : 0;
        float acc[TQ];
#pragma unroll
        for (int j = 0; j < TQ; ++j) acc[j] = 0.f;
#pragma unroll 8
        for (int d = part; d < F; d += nsplit) {
            const float kv = __ldg(kbase + (long long)d * T + tk);
            const float4* q4 = reinterpret_cast<const float4*>(qs + d * TQ);
#pragma unroll
            for (int j4 = 0; j4 < TQ / 4; ++j4) {
                const float4 q = q4[j4];
                acc[j4 * 4 + 0] = fmaf(q.x, kv, acc[j4 * 4 + 0]);
                acc[j4 * 4 + 1] = fmaf(q.y, kv, acc[j4 * 4 + 1]);
                acc[j4 * 4 + 2] = fmaf(q.z, kv, acc[j4 * 4 + 2]);
                acc[j4 * 4 + 3] = fmaf(q.w, kv, acc[j4 * 4 + 3]);
            }
        }
        float* dst = part == 0 ? S : red + (size_t)(part - 1) * TQ * Tp;
#pragma unroll
        for (int j = 0; j < TQ; ++j) dst[j * Tp + tk] = acc[j];
    }
    __syncthreads();
    for (int e = tid; e < TQ * Tp; e += ANT) {
        float v = S[e];
        for (int pp = 1; pp < nsplit; ++pp) v += red[(size_t)(pp - 1) * TQ * Tp + e];
        S[e] = v * scale;
    }
    __syncthreads();

    // softmax over keys: one warp per query row
    for (int j = w; j < TQ; j += ANT / 32) {
        float mx = -INFINITY;
        for (int tk = l; tk < T; tk += 32) mx = fmaxf(mx, S[j * Tp + tk]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int tk = l; tk < T; tk += 32) {
            const float e = expf(S[j * Tp + tk] - mx);
            S[j * Tp + tk] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
        for (int tk = l; tk < T; tk += 32) S[j * Tp + tk] *= inv;
    }
    __syncthreads();
    // transpose the probabilities: PT[tk][j]
    for (int tk = tid; tk < T; tk += ANT) {
#pragma unroll
        for (int j4 = 0; j4 < TQ / 4; ++j4)
            *reinterpret_cast<float4*>(PT + tk * TQ + j4 * 4) =
                make_float4(S[(j4 * 4 + 0) * Tp + tk], S[(j4 * 4 + 1) * Tp + tk], S[(j4 * 4 + 2) * Tp + tk], S[(j4 * 4 + 3) * Tp + tk]);
    }

    // out[f, tq] = sum_tk P[tq][tk] * h[f][tk], FCH value rows at a time; warp w sums keys tk = w, w+8, ...
    const float* hbase = h.p + (long long)b * h.sb + (long long)head * h.sc;
    float* obase = out.p + (long long)b * out.sb + (long long)head * out.sc;
    const int rg = l & 7, qg = l >> 3;            // this thread's 4 rows (4*rg..) and 4 queries (4*qg..)
    for (int fc = 0; fc < F; fc += FCH) {
        __syncthreads();                          // PT complete (first pass) / previous pass done with hT and red
        for (int e = tid; e < FCH * T; e += ANT) {
            const int r = e / T, tk = e - r * T;
            const float v = (fc + r < F) ? __ldg(hbase + (long long)(fc + r) * T + tk) : 0.f;
            hT[tk * FCH + (((r >> 2) ^ (tk & 7)) << 2) + (r & 3)] = v;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int tk = w; tk < T; tk += ANT / 32) {
            const float4 hv = *reinterpret_cast<const float4*>(hT + tk * FCH + ((rg ^ (tk & 7)) << 2));
            const float4 pv = *reinterpret_cast<const float4*>(PT + tk * TQ + (qg << 2));
            const float hh[4] = {hv.x, hv.y, hv.z, hv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(hh[i], pp[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(red + w * (FCH * TQ) + (rg * 4 + i) * TQ + qg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __syncthreads();
        for (int e = tid; e < FCH * TQ; e += ANT) {
            float sacc = 0.f;
#pragma unroll
            for (int ww = 0; ww < ANT / 32; ++ww) sacc += red[ww * (FCH * TQ) + e];
            const int r = e / TQ, j = e % TQ;
            if (fc + r < F && t0 + j < T) obase[(long long)(fc + r) * T + t0 + j] = sacc;
        }
    }
}

void launch_attention(const TV& h, const float* qk, const TV& out, cudaStream_t s) {
    const int F = h.F, T = h.T;
    const size_t smem = ((size_t)F * TQ + (((size_t)TQ * (T + 1) + 3) & ~(size_t)3) + (size_t)T * TQ + (size_t)T * FCH + (size_t)(ANT / 32) * FCH * TQ) * sizeof(float);
    if (smem > 227 * 1024) throw CudaError(cudaErrorInvalidValue, "attention tile exceeds shared memory", __FILE__, __LINE__);
    static SmemConfig configured;
    ensure_dyn_smem(attention_kernel, smem, configured);
    dim3 grid((T + TQ - 1) / TQ, h.C, h.B);
    attention_kernel<<<grid, ANT, smem, s>>>(h, qk, out, 1.0f / sqrtf((float)F));
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
