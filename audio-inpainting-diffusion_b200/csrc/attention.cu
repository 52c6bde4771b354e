// Fused time-attention core (unet.py:353-374): per (clip, head)
//   q[t, d] = qk[b, head*2F + d, t],  k[t, d] = qk[b, head*2F + F + d, t],  v[t, f] = h[b, head, f, t]
//   out[b, head, f, t] = sum_tk softmax_tk(F^-0.5 * q[t,:].k[tk,:]) * v[tk, f]
// One CTA handles TQ queries of one (clip, head): scores live in shared memory (T <= ~1000 frames).
// fp32 CUDA-core version (exact-parity path; < 0.1 % of the forward's FLOPs).
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

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

// ---------------------------------------------------------------------------------------------------------------------------
// tcgen05 version (conv_mode 2): one CTA per (clip, head, tile of 128 queries).
//   S[128, T]  = Q K^T     tcgen05.mma kind::f16, operands split hi + lo in fp16 (q_hi k_hi + q_lo k_hi + q_hi k_lo: fp32-grade
//                          logits), accumulated in TMEM over the feature dimension in chunks of 64
//   P          = exp(scale (S - rowmax))   read from TMEM by the row's thread (tcgen05.ld), written to shared memory as the fp16
//                          A operand of the second product; the row sums stay in registers
//   O[128, F]  = P V       tcgen05.mma, V = h^T staged 64 value rows at a time; O reuses the TMEM columns of S
//   out        = O / rowsum, stored from TMEM lanes (one query per thread, coalesced along t for every value row)
// Operand layout in shared memory: K-major SWIZZLE_NONE canonical form, [8-element k chunk][row][8] fp16 (rows 16 bytes apart,
// SBO = 128, LBO = rows * 16), filled by the CTA's threads (the sources are fp32 [d][t] / [f][t] planes, so a bulk copy could not
// convert or transpose): a thread loads 8 features of one query / key (each load coalesced across the warp) and stores one
// 16-byte row.  T <= 256 keys (N of the first product), F <= 512.
static constexpr int ATC_THREADS = 256;
static constexpr int ATC_DC = 64;      // features per stage of Q K^T
static constexpr int ATC_NF = 64;      // value rows per stage of P V

__device__ __forceinline__ void split_store8(const float* v, uint4* hi, uint4* lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half a = __float2half_rn(v[2 * i]), b = __float2half_rn(v[2 * i + 1]);
        h[i] = __halves2half2(a, b);
        if (lo) l[i] = __halves2half2(__float2half_rn(v[2 * i] - __half2float(a)), __float2half_rn(v[2 * i + 1] - __half2float(b)));
    }
    *hi = *reinterpret_cast<uint4*>(h);
    if (lo) *lo = *reinterpret_cast<uint4*>(l);
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(TV h, const float* __restrict__ qk, TV out, float scale, int Tk16) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* smem = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);
    const int F = h.F, T = h.T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int head = blockIdx.y, b = blockIdx.z, t0 = blockIdx.x * 128;
    const float* qbase = qk + ((long long)b * h.C * 2 * F + (long long)head * 2 * F) * T;
    const float* kbase = qbase + (long long)F * T;
    const float* hbase = h.p + (long long)b * h.sb + (long long)head * h.sc;
    float* obase = out.p + (long long)b * out.sb + (long long)head * out.sc;
    // shared memory map (bytes): phase 1: Qhi | Qlo (8 planes x 128 rows x 16) | Khi | Klo (8 planes x Tk16 rows x 16);
    //                            phase 2: P (Tk16/8 planes x 128 rows x 16) | V (Tk16/8 planes x 64 rows x 16)
    const uint32_t planeQ = 128u * 16u, planeK = (uint32_t)Tk16 * 16u, planeV = (uint32_t)ATC_NF * 16u;
    uint8_t* Qhi = smem; uint8_t* Qlo = Qhi + 8 * planeQ; uint8_t* Khi = Qlo + 8 * planeQ; uint8_t* Klo = Khi + 8 * planeK;
    uint8_t* Ps = smem; uint8_t* Vs = Ps + (size_t)(Tk16 / 8) * planeQ;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;

    // ---------------- S = Q K^T ----------------
    const uint32_t idesc_s = (1u << 4) | ((uint32_t)(Tk16 >> 3) << 17) | ((128u >> 4) << 24);
    for (int d0 = 0; d0 < F; d0 += ATC_DC) {
        const int np = min(ATC_DC, F - d0) / 8;            // feature planes in this stage (F is a multiple of 8... checked by the launcher)
        // queries: 128 rows x np planes; keys: Tk16 rows x np planes.  item = (plane, row), rows fastest across lanes
        for (int it = tid; it < np * 128; it += ATC_THREADS) {
            const int pl = it >> 7, r = it & 127, t = t0 + r;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = t < T ? __ldg(qbase + (long long)(d0 + pl * 8 + j) * T + t) : 0.f;
            split_store8(v, reinterpret_cast<uint4*>(Qhi + pl * planeQ + r * 16), reinterpret_cast<uint4*>(Qlo + pl * planeQ + r * 16));
        }
        for (int it = tid; it < np * Tk16; it += ATC_THREADS) {
            const int pl = it / Tk16, r = it - pl * Tk16;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = r < T ? __ldg(kbase + (long long)(d0 + pl * 8 + j) * T + r) : 0.f;
            split_store8(v, reinterpret_cast<uint4*>(Khi + pl * planeK + r * 16), reinterpret_cast<uint4*>(Klo + pl * planeK + r * 16));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < np / 2; ++ks) {          // 16 features per MMA = 2 planes
                const uint64_t ah = make_desc(smem_u32(Qhi + 2 * ks * planeQ), planeQ, 128), al = make_desc(smem_u32(Qlo + 2 * ks * planeQ), planeQ, 128);
                const uint64_t bh = make_desc(smem_u32(Khi + 2 * ks * planeK), planeK, 128), bl = make_desc(smem_u32(Klo + 2 * ks * planeK), planeK, 128);
                tc_mma_f16(tmem, ah, bh, idesc_s, (d0 > 0 || ks > 0) ? 1u : 0u);
                tc_mma_f16(tmem, al, bh, idesc_s, 1u);
                tc_mma_f16(tmem, ah, bl, idesc_s, 1u);
            }
            tc_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;                 // the stage's operands are free again, S is up to date
    }
    tc_fence_after();

    // ---------------- P = exp(scale (S - max)), row sums ----------------
    float rsum = 1.f;
    if (warp < 4) {
        const int r = warp * 32 + lane;                    // TMEM lane = query row
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        float mx = -INFINITY;
        for (int c0 = 0; c0 < Tk16; c0 += 32) {
            uint32_t a[32];
            tmem_ld32_nowait(trow + (uint32_t)c0, a);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c0 + j < T) mx = fmaxf(mx, __uint_as_float(a[j]));
        }
        float sum = 0.f;
        for (int c0 = 0; c0 < Tk16; c0 += 32) {
            uint32_t a[32];
            tmem_ld32_nowait(trow + (uint32_t)c0, a);
            tmem_wait_ld();
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                if (c0 + j8 * 8 >= Tk16) break;
                float e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + j8 * 8 + j;
                    e[j] = c < T ? expf(scale * (__uint_as_float(a[j8 * 8 + j]) - mx)) : 0.f;
                    e[j] = __half2float(__float2half_rn(e[j]));      // the value the second product will see: keep the sum consistent
                    sum += e[j];
                }
                split_store8(e, reinterpret_cast<uint4*>(Ps + (size_t)((c0 >> 3) + j8) * planeQ + r * 16), nullptr);
            }
        }
        rsum = sum;
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // ---------------- O = P V, value rows in stages of 64; out = O / rowsum ----------------
    const uint32_t idesc_o = (1u << 4) | ((uint32_t)(ATC_NF >> 3) << 17) | ((128u >> 4) << 24);
    const int npk = Tk16 / 8;                               // key planes
    for (int f0 = 0; f0 < F; f0 += ATC_NF) {
        // V stage: rows = value rows f0 .. f0+63 (zero beyond F), K = keys.  item = (plane, row), rows fastest across lanes
        for (int it = tid; it < npk * ATC_NF; it += ATC_THREADS) {
            const int pl = it / ATC_NF, r = it - pl * ATC_NF, f = f0 + r;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int tk = pl * 8 + j; v[j] = (f < F && tk < T) ? __ldg(hbase + (long long)f * T + tk) : 0.f; }
            split_store8(v, reinterpret_cast<uint4*>(Vs + pl * planeV + r * 16), nullptr);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < npk / 2; ++ks) {
                const uint64_t a = make_desc(smem_u32(Ps + (size_t)2 * ks * planeQ), planeQ, 128);
                const uint64_t bd = make_desc(smem_u32(Vs + (size_t)2 * ks * planeV), planeV, 128);
                tc_mma_f16(tmem + (uint32_t)f0, a, bd, idesc_o, ks > 0 ? 1u : 0u);
            }
            tc_commit(&bar);
        }
        mbar_wait(&bar, phase); phase ^= 1;
        tc_fence_after();
        if (warp < 4) {
            const int t = t0 + warp * 32 + lane;
            const float inv = 1.f / rsum;
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)f0;
#pragma unroll
            for (int c0 = 0; c0 < ATC_NF; c0 += 32) {
                uint32_t a[32];
                tmem_ld32_nowait(trow + (uint32_t)c0, a);
                tmem_wait_ld();
                if (t < T) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (f0 + c0 + j < F) obase[(long long)(f0 + c0 + j) * T + t] = __uint_as_float(a[j]) * inv;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

bool attention_tc_supported(int F, int T) { return F % 16 == 0 && F <= 512 && T >= 1 && T <= 256; }

void launch_attention_tc(const TV& h, const float* qk, const TV& out, cudaStream_t s) {
    const int F = h.F, T = h.T;
    const int Tk16 = (T + 15) & ~15;
    const size_t p1 = (size_t)16 * 128 * 16 + (size_t)16 * Tk16 * 16;                       // Q hi|lo + K hi|lo, 8 planes each
    const size_t p2 = (size_t)(Tk16 / 8) * 128 * 16 + (size_t)(Tk16 / 8) * ATC_NF * 16;      // P + one V stage
    const size_t smem = std::max(p1, p2) + 128;
    static SmemConfig configured;
    ensure_dyn_smem(attention_tc_kernel, smem, configured);
    dim3 grid((T + 127) / 128, h.C, h.B);
    attention_tc_kernel<<<grid, ATC_THREADS, smem, s>>>(h, qk, out, 1.0f / sqrtf((float)F), Tk16);
    AID_COUNT_LAUNCH(1);
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
