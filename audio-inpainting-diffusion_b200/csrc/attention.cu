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

__global__ void __launch_bounds__(ANT)
attention_kernel(TV h, const float* __restrict__ qk, TV out, float scale) {
    extern __shared__ float sm[];
    const int F = h.F, T = h.T;
    const int Tp = T + 1;
    float* qs = sm;                    // [F][TQ]
    float* S = qs + (size_t)F * TQ;    // [TQ][Tp]
    float* hs = S + (size_t)TQ * Tp;   // [FCH][Tp]
    float* os = hs + (size_t)FCH * Tp; // [FCH][TQ]

    const int tid = threadIdx.x;
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

    // scores: thread <-> key column(s)
    for (int tk = tid; tk < T; tk += ANT) {
        float acc[TQ];
#pragma unroll
        for (int j = 0; j < TQ; ++j) acc[j] = 0.f;
        for (int d = 0; d < F; ++d) {
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
#pragma unroll
        for (int j = 0; j < TQ; ++j) S[j * Tp + tk] = acc[j] * scale;
    }
    __syncthreads();

    // softmax over keys: one warp per query row
    {
        const int w = tid >> 5, l = tid & 31;
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
    }
    __syncthreads();

    // out[f, tq] = sum_tk P[tq][tk] * h[f][tk], FCH value rows at a time
    const float* hbase = h.p + (long long)b * h.sb + (long long)head * h.sc;
    float* obase = out.p + (long long)b * out.sb + (long long)head * out.sc;
    for (int fc = 0; fc < F; fc += FCH) {
        for (int e = tid; e < FCH * T; e += ANT) {
            const int r = e / T, tk = e % T;
            hs[r * Tp + tk] = (fc + r < F) ? __ldg(hbase + (long long)(fc + r) * T + tk) : 0.f;
        }
        __syncthreads();
        {
            const int r = tid & 31, jq = tid >> 5;  // rows x (queries jq, jq+8)
            float a0 = 0.f, a1 = 0.f;
            const float* hr = hs + r * Tp;
            const float* p0 = S + jq * Tp;
            const float* p1 = S + (jq + 8) * Tp;
            for (int tk = 0; tk < T; ++tk) {
                const float hv = hr[tk];
                a0 = fmaf(p0[tk], hv, a0);
                a1 = fmaf(p1[tk], hv, a1);
            }
            os[r * TQ + jq] = a0;
            os[r * TQ + jq + 8] = a1;
        }
        __syncthreads();
        for (int e = tid; e < FCH * TQ; e += ANT) {
            const int r = e / TQ, j = e % TQ;
            if (fc + r < F && t0 + j < T) obase[(long long)(fc + r) * T + t0 + j] = os[e];
        }
        // next iteration's hs/os writes are ordered by the __syncthreads above and below
        __syncthreads();
    }
}

void launch_attention(const TV& h, const float* qk, const TV& out, cudaStream_t s) {
    const int F = h.F, T = h.T;
    const size_t smem = ((size_t)F * TQ + (size_t)TQ * (T + 1) + (size_t)FCH * (T + 1) + (size_t)FCH * TQ) * sizeof(float);
    if (smem > 227 * 1024) throw CudaError(cudaErrorInvalidValue, "attention tile exceeds shared memory", __FILE__, __LINE__);
    static size_t configured = 0;
    if (smem > configured) {
        AID_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((T + TQ - 1) / TQ, h.C, h.B);
    attention_kernel<<<grid, ANT, smem, s>>>(h, qk, out, 1.0f / sqrtf((float)F));
    AID_COUNT_LAUNCH(1);
}

}  // namespace aid
