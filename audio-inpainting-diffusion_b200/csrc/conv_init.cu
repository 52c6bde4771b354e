// Fused init block of an encoder level (unet.py:412-415, 452-493 with dim = 2, 1x1 kernels, one layer; called at unet.py:673-675):
//     y   = proj_in(x2)                                   x2: the level's two CQT channels (re, im) [B, 2, F, T], y: [B, N, F, T]
//     x   = (y + gate * H(GELU(GroupNorm8(y) * gamma * (1 + affine)))) / sqrt 2          H: 1x1, N -> N
//     out = (x + res_conv(x2)) / sqrt 2
// Un-fused this is five launches that move 28 bytes per output element (write y, read y / write the operand, read operand + y / write
// x, read x / write out).  Everything except the N x N contraction is a function of the TWO input values of the pixel, so one kernel
// reads x2 (8 bytes per pixel) and writes out:
//   * the group-norm statistics of y are those of a linear map of x2: per clip they follow from the five second moments of x2
//     (init_prep_kernel), no pass over y is needed -- y is never materialised;
//   * transform warps compute y from x2, normalise, GELU, convert and write the K-major SWIZZLE_128B operand rows into a two-slot ring;
//   * one elected thread issues the N / 16 MMAs of a 128-pixel unit (weights resident in shared memory, accumulators in TMEM);
//   * the epilogue (tc_epilogue.cuh, RES = 1) rebuilds y and res_conv(x2) from x2 with two per-clip coefficient tables:
//         out = acc * (gate / 2) + (x2[0] * c0 + x2[1] * c1) / 2,     c = w_in + sqrt 2 * w_res,
//     and accumulates the statistics of out for the next block.
// Channels: 64, 96, 128 (levels 0-4 of the paper network; 256 would need 128 KB of weights next to a 128 KB operand ring).
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_epilogue.cuh"

namespace aid {

static constexpr int IB_EPI = 8, IB_TR = 8, IB_WARP_W = 16, IB_WARP_MMA = 17, IB_THREADS = 18 * 32;
static constexpr int IB_NACC = 4, IB_ACC_STRIDE = 128, IB_SLOTS = 2;
static constexpr int IB_BAR_BYTES = 256, IB_GATE_BYTES = 8 * 256 * 4, IB_STAT_BYTES = 16384, IB_TTAB_BYTES = 8 * 64 * 4;

struct InitArgs {
    TV x2, out;                       // [B, 2, F, T] input, [B, N, F, T] output view
    const float* ttab;                // [B][N][4] = {w_in0, w_in1, cu, ch}: y = w_in0 x2[0] + w_in1 x2[1]; operand = 16 GELU(y s) with cu = |s| sqrt(log2 e / 2), ch = 8 s
    const float* c0tab; const float* c1tab;   // [B][N] residual coefficients of the epilogue
    const float* gate; long long gate_bstride;
    const __half* w;                  // H weights, launch_pack_weight_tc2 layout for a 1x1 convolution: [G][N][64]
    double* stats_out;
    int B, N, F, T, tiles_t, n_units;
};

struct InitUnit { int b, f, t0; };
__device__ __forceinline__ InitUnit init_unit(const InitArgs& p, int u) {
    InitUnit i;
    const int tt = u % p.tiles_t, bf = u / p.tiles_t;
    i.f = bf % p.F; i.b = bf / p.F; i.t0 = tt * 128;
    return i;
}
struct InitUnitIter {
    const InitArgs& p; int pofs, step, u, ab = 0; uint32_t aph = 0;
    __device__ InitUnitIter(const InitArgs& p_, int u0, int step_, int pofs_) : p(p_), pofs(pofs_), step(step_), u(u0) {}
    __device__ __forceinline__ bool next(EpiUnit& d) {
        if (u >= p.n_units) return false;
        const InitUnit iu = init_unit(p, u);
        d.b = iu.b; d.nt = 0; d.ok = true;
        d.pix = (long long)iu.f * p.T + iu.t0 + pofs;
        d.tcol = (uint32_t)(ab * IB_ACC_STRIDE); d.ab = ab; d.aph = aph; d.first = true; d.last = true;
        if (++ab == IB_NACC) { ab = 0; aph ^= 1; }
        u += step;
        return true;
    }
};

// N: channels.  G = ceil(N / 64) operand groups of 64 channels (the last one half empty for N = 96: its k-steps 2, 3 are not issued).
template <int N>
__global__ void __launch_bounds__(IB_THREADS, 1) init_block_kernel(const __grid_constant__ InitArgs p) {
    constexpr int G = (N + 63) / 64, NCH = N / 8;                 // operand groups, 8-channel chunks
    constexpr int SLOT = G * 16384, WBYTES = G * N * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem;
    uint8_t* wsm = ring + (size_t)IB_SLOTS * SLOT;
    uint8_t* bar_base = wsm + WBYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
    uint64_t* a_empty = a_full + IB_SLOTS;
    uint64_t* tmem_full = a_empty + IB_SLOTS;
    uint64_t* tmem_empty = tmem_full + IB_NACC;
    uint64_t* w_full = tmem_empty + IB_NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* gsm_base = reinterpret_cast<float*>(bar_base + IB_BAR_BYTES);
    double* sacc_base = reinterpret_cast<double*>(bar_base + IB_BAR_BYTES + IB_GATE_BYTES);
    float* ttab_sm = reinterpret_cast<float*>(bar_base + IB_BAR_BYTES + IB_GATE_BYTES + IB_STAT_BYTES);   // [8 warps][16 channels][4]
    const int u0 = blockIdx.x, ustep = gridDim.x;

    if (warp == IB_WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < IB_SLOTS; ++s) { mbar_init(a_full + s, IB_TR); mbar_init(a_empty + s, 1); }
            for (int s = 0; s < IB_NACC; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, IB_EPI); }
            mbar_init(w_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == IB_WARP_W) {
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)WBYTES);
            for (int g = 0; g < G; ++g) bulk_g2s(wsm + g * (N * 128), p.w + (size_t)g * N * 64, N * 128, w_full);
        }
        __syncwarp();
    } else if (warp >= IB_EPI && warp < IB_EPI + IB_TR) {
        // ===================== transform: warp w = operand chunks w and w + 8 (8 channels each), lane l = pixels l, l+32, l+64, l+96 =====================
        const int w = warp - IB_EPI;
        constexpr int NMY = NCH > 8 ? 2 : 1;                  // chunk slots of a warp (the second one exists for w + 8 < NCH)
        float* tt = ttab_sm + w * 64;                          // this warp's constants: [slot][8 channels][4]
        const uint32_t ring_u = smem_u32(ring);
        int n = 0, b_cur = -1;
        for (int u = u0; u < p.n_units; u += ustep, ++n) {
            const InitUnit iu = init_unit(p, u);
            if (iu.b != b_cur) {
                b_cur = iu.b;
                __syncwarp();
                if (lane < 8 * NMY) {
                    const int ch = 8 * (w + 8 * (lane >> 3)) + (lane & 7);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ch < N) v = __ldg(reinterpret_cast<const float4*>(p.ttab + ((long long)iu.b * N + ch) * 4));
                    *reinterpret_cast<float4*>(tt + lane * 4) = v;
                }
                __syncwarp();
            }
            const float* q = p.x2.p + (long long)iu.b * p.x2.sb + (long long)iu.f * p.T + iu.t0 + lane;
            float x0[4], x1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { x0[i] = __ldg(q + i * 32); x1[i] = __ldg(q + p.x2.sc + i * 32); }
            const int slot = n % IB_SLOTS;
            bool waited = false;
#pragma unroll
            for (int cs = 0; cs < NMY; ++cs) {
                const int chunk = w + 8 * cs;
                if (chunk >= NCH) break;
                uint32_t hp[4][4];
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const float4 ca = *reinterpret_cast<const float4*>(tt + (cs * 8 + 2 * pr) * 4);        // channel 2 pr    : w0, w1, cu, ch
                    const float4 cb = *reinterpret_cast<const float4*>(tt + (cs * 8 + 2 * pr + 1) * 4);    // channel 2 pr + 1
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float ya = fmaf(ca.x, x0[i], ca.y * x1[i]), yb = fmaf(cb.x, x0[i], cb.y * x1[i]);
                        const float2 r2 = gelu16_tc2_folded2(make_float2(ya, yb), make_float2(ca.z, cb.z), make_float2(ca.w, cb.w));
                        hp[i][pr] = pack_half2_sat(r2.x, r2.y);
                    }
                }
                if (!waited) { if (n >= IB_SLOTS) mbar_wait(a_empty + slot, (uint32_t)((n / IB_SLOTS) - 1) & 1u); waited = true; }
                const uint32_t gbase = ring_u + (uint32_t)slot * SLOT + (uint32_t)(chunk >> 3) * 16384u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t row = (uint32_t)(i * 32 + lane);
                    const uint32_t addr = gbase + row * 128u + ((((uint32_t)chunk & 7u) ^ (row & 7u)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[i][0]), "r"(hp[i][1]), "r"(hp[i][2]), "r"(hp[i][3]) : "memory");
                }
            }
            if (!waited && n >= IB_SLOTS) mbar_wait(a_empty + slot, (uint32_t)((n / IB_SLOTS) - 1) & 1u);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + slot);
        }
    } else if (warp == IB_WARP_MMA) {
        if (elect_one_sync()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t adesc = desc_lo_sw128(smem_u32(ring)), bdesc = desc_lo_sw128(smem_u32(wsm));
            mbar_wait(w_full, 0);
            int n = 0, ab = 0; uint32_t aph = 0;
            for (int u = u0; u < p.n_units; u += ustep, ++n) {
                const int slot = n % IB_SLOTS;
                mbar_wait(tmem_empty + ab, aph ^ 1);
                mbar_wait(a_full + slot, (uint32_t)(n / IB_SLOTS) & 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(ab * IB_ACC_STRIDE);
                const uint32_t a = adesc + (uint32_t)slot * (SLOT >> 4);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    if (g + 1 < G || N % 64 == 0) tc_mma_k<1, 4>(d, a + (uint32_t)g * (16384u >> 4), bdesc + (uint32_t)g * ((N * 128) >> 4), idesc, g ? 1u : 0u);
                    else tc_mma_k<1, 2>(d, a + (uint32_t)g * (16384u >> 4), bdesc + (uint32_t)g * ((N * 128) >> 4), idesc, g ? 1u : 0u);
                }
                tc_commit(a_empty + slot);
                tc_commit(tmem_full + ab);
                if (++ab == IB_NACC) { ab = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp < IB_EPI) {
        EpiArgs ea{p.out, TV(), p.gate, p.gate_bstride, 0.5f, p.stats_out, N, 1};
        ea.x2 = p.x2; ea.c0tab = p.c0tab; ea.c1tab = p.c1tab; ea.tab_bstride = N;
        InitUnitIter it(p, u0, ustep, (warp & 3) * 32 + lane);
        if constexpr (N == 64) epilogue_fast<false, 16, 2, 8, InitUnitIter, IB_EPI, 1, 256>(ea, it, warp, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty);
        else if constexpr (N == 96) epilogue_fast<false, 12, 4, 12, InitUnitIter, IB_EPI, 1, 256>(ea, it, warp, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty);
        else epilogue_fast<false, 16, 4, 16, InitUnitIter, IB_EPI, 1, 256>(ea, it, warp, lane, tmem_base, gsm_base, sacc_base, tmem_full, tmem_empty);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == IB_WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Per clip: second moments of x2 -> group statistics of y = W_in x2 -> the transform's constants and the epilogue's coefficient tables.
// init_moments_kernel: IB_MSEG blocks per clip, each reduces a contiguous segment (fixed-order double tree); init_prep_kernel: one block
// per clip adds the segment sums in ascending order -- deterministic and independent of the batch.
static constexpr int IB_MSEG = 64;
__global__ void __launch_bounds__(256) init_moments_kernel(TV x2, double* __restrict__ part) {     // part: [B][IB_MSEG][5]
    const int b = blockIdx.y, P = x2.F * x2.T;
    const int seg = (P + IB_MSEG - 1) / IB_MSEG, e0 = blockIdx.x * seg, e1 = min(P, e0 + seg);
    const float* p0 = x2.p + (long long)b * x2.sb;
    const float* p1 = p0 + x2.sc;
    double m[5] = {0, 0, 0, 0, 0};
#pragma unroll 4
    for (int e = e0 + threadIdx.x; e < e1; e += 256) {
        const double a = (double)__ldg(p0 + e), c = (double)__ldg(p1 + e);
        m[0] += a; m[1] += c; m[2] += a * a; m[3] += a * c; m[4] += c * c;
    }
    __shared__ double red[5][256];
#pragma unroll
    for (int k = 0; k < 5; ++k) red[k][threadIdx.x] = m[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
#pragma unroll
            for (int k = 0; k < 5; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 5) part[((long long)b * IB_MSEG + blockIdx.x) * 5 + threadIdx.x] = red[threadIdx.x][0];
}

__global__ void __launch_bounds__(128) init_prep_kernel(const double* __restrict__ part, const float* __restrict__ w_in, const float* __restrict__ w_res,
                                                        const float* __restrict__ gamma, const float* __restrict__ affine, long long affine_bstride, int N,
                                                        double npg, float* __restrict__ ttab, float* __restrict__ c0tab, float* __restrict__ c1tab) {
    const int b = blockIdx.x;
    __shared__ double M[5];
    __shared__ float inv[8];
    if (threadIdx.x < 5) {
        double v = 0;
        for (int s = 0; s < IB_MSEG; ++s) v += part[((long long)b * IB_MSEG + s) * 5 + threadIdx.x];
        M[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {      // group g: channels [g N / 8, (g + 1) N / 8); unbiased std of y over the group, as BiasFreeGroupNorm
        const int gcn = N / 8;
        double s1 = 0, s2 = 0;
        for (int c = threadIdx.x * gcn; c < (threadIdx.x + 1) * gcn; ++c) {
            const double a = (double)w_in[c], d = (double)w_in[N + c];
            s1 += a * M[0] + d * M[1];
            s2 += a * a * M[2] + 2.0 * a * d * M[3] + d * d * M[4];
        }
        double var = (s2 - s1 * s1 / npg) / (npg - 1.0);
        var = var > 0.0 ? var : 0.0;
        inv[threadIdx.x] = 1.f / ((float)sqrt(var) + 1e-7f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += 128) {
        const float mod = affine ? (1.f + affine[(long long)b * affine_bstride + c]) : 1.f;
        const float s = gamma[c] * mod * inv[c / (N / 8)];
        float* t = ttab + ((long long)b * N + c) * 4;
        t[0] = w_in[c]; t[1] = w_in[N + c]; t[2] = fabsf(s) * 0.84932180028801904272f; t[3] = 8.f * s;
        c0tab[(long long)b * N + c] = fmaf(1.41421356237309504880f, w_res[c], w_in[c]);
        c1tab[(long long)b * N + c] = fmaf(1.41421356237309504880f, w_res[N + c], w_in[N + c]);
    }
}

bool init_block_supported(int N, int T) { return (N == 64 || N == 96 || N == 128) && T % 128 == 0; }
size_t init_block_scratch_floats(int B, int N) { return (size_t)B * N * 6 + (size_t)B * IB_MSEG * 5 * 2; }   // ttab (4) + c0tab + c1tab + segment moments (doubles)

template <int N>
static void launch_init_block_n(const InitArgs& p, int num_sms, cudaStream_t s) {
    constexpr int G = (N + 63) / 64;
    constexpr size_t smem = 1024 + (size_t)IB_SLOTS * G * 16384 + (size_t)G * N * 128 + IB_BAR_BYTES + IB_GATE_BYTES + IB_STAT_BYTES + IB_TTAB_BYTES;
    static SmemConfig configured;
    ensure_dyn_smem(init_block_kernel<N>, smem, configured);
    init_block_kernel<N><<<std::min(p.n_units, num_sms), IB_THREADS, smem, s>>>(p);
}

// w_in, w_res: K-major 1x1 weights [2][N] (pack_conv_weight layout); wH: launch_pack_weight_tc2 packing of the N x N 1x1 convolution;
// gamma [N]; affine / gate: per-clip vectors with stride *_bstride (0: shared); scratch: init_block_scratch_floats(B, N) floats
void launch_init_block(const TV& x2, const float* w_in, const float* w_res, const __half* wH, const float* gamma, const float* affine,
                       long long affine_bstride, const float* gate, long long gate_bstride, const TV& out, double* stats_out, float* scratch,
                       int num_sms, cudaStream_t s) {
    const int N = out.C, B = x2.B;
    if (!init_block_supported(N, x2.T) || x2.C != 2 || x2.F != out.F || x2.T != out.T) throw CudaError(cudaErrorInvalidValue, "init_block: unsupported shape", __FILE__, __LINE__);
    float* ttab = scratch; float* c0 = scratch + (size_t)B * N * 4; float* c1 = c0 + (size_t)B * N;
    const double npg = (double)(N / 8) * x2.F * x2.T;
    double* part = reinterpret_cast<double*>(c1 + (size_t)B * N);      // (B * N * 6 floats: a multiple of 8 bytes, N % 8 == 0)
    init_moments_kernel<<<dim3(IB_MSEG, B), 256, 0, s>>>(x2, part);
    init_prep_kernel<<<B, 128, 0, s>>>(part, w_in, w_res, gamma, affine, affine_bstride, N, npg, ttab, c0, c1);
    InitArgs p{};
    p.x2 = x2; p.out = out; p.ttab = ttab; p.c0tab = c0; p.c1tab = c1; p.gate = gate; p.gate_bstride = gate_bstride; p.w = wH; p.stats_out = stats_out;
    p.B = B; p.N = N; p.F = x2.F; p.T = x2.T; p.tiles_t = x2.T / 128; p.n_units = B * x2.F * p.tiles_t;
    if (N == 64) launch_init_block_n<64>(p, num_sms, s);
    else if (N == 96) launch_init_block_n<96>(p, num_sms, s);
    else launch_init_block_n<128>(p, num_sms, s);
    AID_COUNT_LAUNCH(3);
}

}  // namespace aid
