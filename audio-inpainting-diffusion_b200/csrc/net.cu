// Network executor + C ABI (include/aid_b200.h) for the CQT-octave U-Net denoiser.
//
// Host-side mirror of Unet_CQT_oct_with_attention (unet.py:583-845): the constructor builds the same
// module tree / state-dict schema, aid_unet_forward() replays forward() as a fixed sequence of kernel
// launches on one stream over a caller-owned workspace (first-fit planned, no allocation per call).
// Concatenations (unet.py:769-774, 814) and slices (unet.py:821-822) are strided views, never copies.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/aid_b200.h"
#include "common.cuh"
#include "cqt_plan.hpp"

namespace aid {

unsigned long long g_launch_count = 0;

int device_sm_count() {
    static int sms[kMaxDevices] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        sms[dev] = v;
    }
    return sms[dev];
}

static const float kInvSqrt2 = 0.70710678118654752440f;

// ---- weights --------------------------------------------------------------------------------------
struct Weight {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> host;
    bool loaded = false;
    bool ignored = false;  // accepted but unused (resampler kernels are constants)
    size_t numel() const { size_t n = 1; for (auto d : shape) n *= (size_t)d; return n; }
};

struct ConvW {
    int widx = -1; float* wp = nullptr; __half* wtc = nullptr; int Cin = 0, Cout = 0, KF = 1, KT = 1;
    float* wpT = nullptr;   // K-major weights of the transposed (data-gradient) convolution, built on the first VJP call
    __half* wtcT = nullptr; // conv_mode 2: the same for the tcgen05 path (dilated 5x3 layers only)
    __half* wcomb = nullptr; // conv_mode 2: packing of the fused layer kernel (conv_comb.cu) where it differs from wtc (96-channel 5x3 layers)
};
struct LinRef { int w = -1, b = -1; int off = -1, N = 0; };
struct NormRef { int widx = -1; float* gamma = nullptr; };

struct ResBlk {
    int dim = 0, dim_out = 0, N = 0, nd = 0, Fdim = 0;
    bool k1x1 = false, after = false, attn = false;
    ConvW proj_in, res_conv, proj_out, a_in, a_out, qk;
    std::vector<ConvW> H;
    std::vector<LinRef> affine, gate;
    std::vector<NormRef> norm;
    LinRef affine2, gate2;
    NormRef norm2;
};

struct Level { ResBlk init, main; ConvW pyr; };

// What a ResnetBlock's backward needs from its forward (views into the caller's workspace, kept alive by the taped forward)
struct BlkTape {
    TV in, x0;            // block input; after proj_in (== in when the block has none); x0.stats valid
    TV h, qk, x1;         // attention sub-block: projected heads [B,8,F,T], q|k [B,16F,1,T], its output (== x0 without attention)
    std::vector<TV> xs;   // xs[i] = input of dilated layer i (stats valid)
};
struct Tape {
    bool valid = false;
    int B = 0, nsig = 1;
    float in_scale = 1.f, out_scale = 1.f, skip_scale = 0.f;
    float* mod = nullptr;
    std::vector<BlkTape> init, main, ups_main, ups_out;
    BlkTape mid_main, mid_out;
    float2 *spec = nullptr, *tmp = nullptr, *fscr = nullptr, *Y = nullptr;
    char* ws = nullptr; size_t ws_bytes = 0;
    unsigned long long generation = 0;
};

// ---- workspace arena: first-fit over a caller-owned slab; "dry" mode only measures the peak -----------
struct Arena {
    char* base = nullptr; size_t cap = 0, peak = 0; bool dry = true;
    std::map<size_t, size_t> used;  // offset -> size
    void reset(char* b, size_t c, bool d) { base = b; cap = c; dry = d; peak = 0; used.clear(); }
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (bytes == 0) bytes = 256;
        size_t off = 0;
        for (auto& kv : used) {
            if (kv.first >= off + bytes) break;
            off = std::max(off, kv.first + kv.second);
        }
        used[off] = bytes;
        peak = std::max(peak, off + bytes);
        if (!dry && off + bytes > cap) throw std::runtime_error("workspace too small");
        return dry ? reinterpret_cast<void*>(0x100000 + off) : base + off;
    }
    void release(void* p) {
        const size_t off = dry ? reinterpret_cast<size_t>(p) - 0x100000 : (size_t)((char*)p - base);
        used.erase(off);
    }
};

struct Net {
    aid_config cfg{};
    int device = 0;
    std::string err;
    std::vector<Weight> weights;
    std::map<std::string, int> windex;
    std::vector<Level> downs;
    ResBlk mid_main, mid_out;
    std::vector<ResBlk> ups_main, ups_out;
    std::vector<LinRef*> lin_all;  // every adaLN Linear, in mod-vector order
    int total_mod = 0;
    int emb_idx[7] = {-1, -1, -1, -1, -1, -1, -1};  // RFF_freq, W0,b0,W1,b1,W2,b2
    bool finalized = false;
    // device state
    float* dweights = nullptr;   // packed conv weights + gammas + embedding + mod matrix
    float* d_emb[7] = {nullptr};
    float* d_modW = nullptr; float* d_modB = nullptr;
    CqtPlanHost plan;
    CqtTables tabs;
    FftPlan fft;
    void* d_tables = nullptr;
    int n_stat_slots = 0;
    int num_sms = 0;   // multiprocessors of `device`, read at finalize
    int tc_parts() const { return cfg.conv_mode == 1 ? 2 : 1; }  // fp16 parts per tensor-core operand (conv_mode 1: hi + lo)
    __half* dweights_tc = nullptr;  // split-fp16 packed weights of the tcgen05 convolutions (conv_mode 1)
    std::map<std::string, float*> probes;  // debug: name -> caller buffer that receives a contiguous copy
    // optional per-launch timing of the convolution kernels (bench.py roofline): CUDA events on the launching stream
    struct ProfRec { cudaEvent_t e0, e1; int kind; double flops, bytes; };
    bool prof = false;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    // conv_mode 2 saturation accounting (aid_debug_saturation): fp16 operands are x16 / x1024 scaled and converted with
    // satfinite; d_sat[0] counts activation values clamped by the operand passes while counting is enabled, weight_sat the
    // weight values clamped at finalize
    unsigned long long* d_sat = nullptr;
    bool count_sat = false;
    bool fuse_init = true, fuse_comb = true, fuse_out = true, fuse_up = true;   // aid_debug_fusion: run the un-fused twins of conv_init.cu / conv_comb.cu (parity tests)
    unsigned long long weight_sat = 0;
    // input-gradient path (aid_unet_forward_tape / aid_unet_backward)
    Tape tape;
    Arena tape_arena;              // arena state behind the taped forward: the backward keeps allocating in the same workspace
    float* dweights_T = nullptr;   // transposed convolution weights (lazy)
    __half* dweights_tcT = nullptr; // conv_mode 2: fp16 operand-layout weights of the transposed dilated convolutions (lazy)
    float* d_adj_tables = nullptr; // CQT adjoint window tables (lazy)
    CqtTables tabs_adjA, tabs_adjS;   // analysis adjoint: dual := win / M; synthesis adjoint: win := dual * M
    bool vjp_ready = false;
};

static int add_weight(Net& n, const std::string& name, std::vector<int64_t> shape, bool ignored = false) {
    Weight w; w.name = name; w.shape = std::move(shape); w.ignored = ignored;
    n.weights.push_back(std::move(w));
    n.windex[name] = (int)n.weights.size() - 1;
    return (int)n.weights.size() - 1;
}

static ConvW add_conv(Net& n, const std::string& name, int Cout, int Cin, int KF, int KT, bool conv1d = false) {
    ConvW c; c.Cin = Cin; c.Cout = Cout; c.KF = KF; c.KT = KT;
    c.widx = conv1d ? add_weight(n, name + ".weight", {Cout, Cin, 1}) : add_weight(n, name + ".weight", {Cout, Cin, KF, KT});
    return c;
}

static LinRef add_lin(Net& n, const std::string& name, int N) {
    LinRef l; l.N = N;
    l.w = add_weight(n, name + ".weight", {N, 256});
    l.b = add_weight(n, name + ".bias", {N});
    l.off = n.total_mod; n.total_mod += N;
    return l;
}

// unet.py:382-448
static void build_resblk(Net& n, ResBlk& k, const std::string& p, int dim, int dim_out, int nd, bool k1x1, bool after,
                         bool attn, int Fdim) {
    k.dim = dim; k.dim_out = dim_out; k.nd = nd; k.k1x1 = k1x1; k.after = after; k.attn = attn; k.Fdim = Fdim;
    k.N = after ? dim : dim_out;
    const int N = k.N;
    if (N % 8 != 0) throw std::invalid_argument("block width must be a multiple of 8 (8 norm groups): " + p);
    if (after && N != dim_out) k.proj_out = add_conv(n, p + ".proj_out", dim_out, N, 1, 1);
    if (dim != dim_out) k.res_conv = add_conv(n, p + ".res_conv", dim_out, dim, 1, 1);
    if (dim != N) k.proj_in = add_conv(n, p + ".proj_in", N, dim, 1, 1);
    for (int i = 0; i < nd; ++i) {
        const std::string si = std::to_string(i);
        NormRef nr; nr.widx = add_weight(n, p + ".norm." + si + ".gamma", {1, N, 1, 1});
        k.norm.push_back(nr);
        k.affine.push_back(add_lin(n, p + ".affine." + si, N));
        k.gate.push_back(add_lin(n, p + ".gate." + si, N));
        k.H.push_back(k1x1 ? add_conv(n, p + ".H." + si, N, N, 1, 1) : add_conv(n, p + ".H." + si, N, N, 5, 3));
    }
    if (attn) {
        const int heads = n.cfg.num_heads;
        k.norm2.widx = add_weight(n, p + ".norm2.gamma", {1, N, 1, 1});
        k.affine2 = add_lin(n, p + ".affine2", N);
        k.gate2 = add_lin(n, p + ".gate2", N);
        k.qk = add_conv(n, p + ".attn_block.qk", 2 * heads * Fdim, heads * Fdim, 1, 1, /*conv1d=*/true);
        k.a_in = add_conv(n, p + ".attn_block.proj_in", heads, N, 1, 1);
        k.a_out = add_conv(n, p + ".attn_block.proj_out", N, heads, 1, 1);
    }
}

static void collect_lins(Net& n, ResBlk& k) {
    for (auto& l : k.affine) n.lin_all.push_back(&l);
    for (auto& l : k.gate) n.lin_all.push_back(&l);
    if (k.attn) { n.lin_all.push_back(&k.affine2); n.lin_all.push_back(&k.gate2); }
}

// unet.py:587-721
static void build_net(Net& n) {
    const aid_config& c = n.cfg;
    if (c.num_octs < 1 || c.num_octs > AID_MAX_OCTS) throw std::invalid_argument("num_octs out of range");
    if (c.emb_dim != 256) throw std::invalid_argument("emb_dim must be 256");
    if (c.num_bottleneck_layers != 1) throw std::invalid_argument("num_bottleneck_layers must be 1");
    if (c.num_heads < 1) throw std::invalid_argument("num_heads must be >= 1");
    if (c.conv_mode < 0 || c.conv_mode > 2) throw std::invalid_argument("conv_mode must be 0 (fp32 CUDA cores), 1 (tcgen05 split-fp16, 3 MMAs) or 2 (tcgen05 single fp16)");
    const int no = c.num_octs, bins = c.bins_per_oct;
    n.emb_idx[0] = add_weight(n, "embedding.RFF_freq", {1, 32});
    const int dims[4] = {64, 128, 256, 256};
    for (int i = 0; i < 3; ++i) {
        n.emb_idx[1 + 2 * i] = add_weight(n, "embedding.MLP." + std::to_string(i) + ".weight", {dims[i + 1], dims[i]});
        n.emb_idx[2 + 2 * i] = add_weight(n, "embedding.MLP." + std::to_string(i) + ".bias", {dims[i + 1]});
    }
    add_weight(n, "downsamplerT.kernel", {8}, true);
    add_weight(n, "upsamplerT.kernel", {8}, true);
    n.downs.resize(no);
    for (int i = 0; i < no; ++i) {
        const int din = i == 0 ? c.Ns[0] : c.Ns[i - 1], dout = c.Ns[i];
        const std::string p = "downs." + std::to_string(i);
        build_resblk(n, n.downs[i].init, p + ".0", 2, din, 1, true, false, false, bins);
        n.downs[i].pyr = add_conv(n, p + ".1", dout, 2, 5, 3);
        build_resblk(n, n.downs[i].main, p + ".2", din, dout, c.num_dils[i], false, false, c.attention_layers[i] != 0, (i + 1) * bins);
    }
    build_resblk(n, n.mid_out, "middle.0.0", c.Ns[no - 1], 2, 1, true, true, false, no * bins);
    build_resblk(n, n.mid_main, "middle.0.1", c.Ns[no - 1], c.Ns[no - 1], c.num_dils[no - 1], false, false,
                 c.attention_layers[no] != 0, no * bins);
    n.ups_main.resize(no); n.ups_out.resize(no);
    for (int i = 0; i < no; ++i) {
        const int j = no - 1 - i;
        const int din = 2 * c.Ns[j], dout = j == 0 ? c.Ns[0] : c.Ns[j - 1];
        const std::string p = "ups." + std::to_string(i);
        build_resblk(n, n.ups_out[i], p + ".0", dout, 2, 1, true, true, false, (j + 1) * bins);
        build_resblk(n, n.ups_main[i], p + ".1", din, dout, c.num_dils[j], false, false, c.attention_layers[j] != 0, (j + 1) * bins);
    }
    for (auto& l : n.downs) { collect_lins(n, l.init); collect_lins(n, l.main); }
    collect_lins(n, n.mid_out); collect_lins(n, n.mid_main);
    for (int i = 0; i < no; ++i) { collect_lins(n, n.ups_out[i]); collect_lins(n, n.ups_main[i]); }
    n.plan.build(no, bins, c.sample_rate, c.audio_len, c.window_kind, c.beta);
}

// ---- device upload / repack ---------------------------------------------------------------------------
// w[co][ci][tap] -> wp[(ci*taps + tap)*Cout + co]
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int taps) {
    const long long n = (long long)Cout * Cin * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long k = i / Cout;
        wp[i] = w[(long long)co * Cin * taps + k];
    }
}

static void upload_tables(Net& n);

static void finalize_net(Net& n) {
    for (auto& w : n.weights)
        if (!w.loaded && !w.ignored) throw std::runtime_error("missing weight: " + w.name);
    DeviceGuard guard(n.device);
    AID_CUDA_CHECK(cudaDeviceGetAttribute(&n.num_sms, cudaDevAttrMultiProcessorCount, n.device));
    size_t total = 0, max_conv = 0;
    auto al = [](size_t v) { return (v + 63) & ~(size_t)63; };
    std::vector<ConvW*> convs;
    std::vector<NormRef*> norms;
    auto visit = [&](ResBlk& k) {
        for (ConvW* c : {&k.proj_in, &k.res_conv, &k.proj_out, &k.a_in, &k.a_out, &k.qk}) if (c->widx >= 0) convs.push_back(c);
        for (auto& c : k.H) convs.push_back(&c);
        for (auto& nr : k.norm) norms.push_back(&nr);
        if (k.attn) norms.push_back(&k.norm2);
    };
    for (auto& l : n.downs) { visit(l.init); visit(l.main); convs.push_back(&l.pyr); }
    visit(n.mid_main); visit(n.mid_out);
    for (auto& k : n.ups_main) visit(k);
    for (auto& k : n.ups_out) visit(k);
    for (auto* c : convs) { size_t e = n.weights[c->widx].numel(); total += al(e); max_conv = std::max(max_conv, e); }
    for (auto* nr : norms) total += al(n.weights[nr->widx].numel());
    for (int i = 0; i < 7; ++i) total += al(n.weights[n.emb_idx[i]].numel());
    total += al((size_t)n.total_mod * 256) + al((size_t)n.total_mod);
    AID_CUDA_CHECK(cudaMalloc(&n.dweights, total * sizeof(float)));
    float* stage = nullptr;
    AID_CUDA_CHECK(cudaMalloc(&stage, max_conv * sizeof(float)));
    size_t off = 0;
    for (auto* c : convs) {
        Weight& w = n.weights[c->widx];
        const size_t e = w.numel();
        AID_CUDA_CHECK(cudaMemcpy(stage, w.host.data(), e * sizeof(float), cudaMemcpyHostToDevice));
        c->wp = n.dweights + off;
        pack_conv_weight_kernel<<<(int)std::min<size_t>(4096, (e + 255) / 256), 256>>>(stage, c->wp, c->Cout, c->Cin, c->KF * c->KT);
        AID_CUDA_CHECK(cudaGetLastError());
        AID_CUDA_CHECK(cudaDeviceSynchronize());
        off += al(e);
    }
    if (n.cfg.conv_mode >= 1) {
        const int parts = n.tc_parts();
        size_t tc_total = 0;
        auto tc_halves = [&](const ConvW* c) {
            return n.cfg.conv_mode == 2 ? tc2_weight_halves(c->Cout, c->Cin, c->KF, c->KT) : (size_t)parts * n.weights[c->widx].numel();
        };
        auto comb96 = [&](const ConvW* c) { return n.cfg.conv_mode == 2 && c->KF == 5 && c->KT == 3 && c->Cin == 96 && c->Cout == 96; };
        for (auto* c : convs) if (conv_tc_supported(c->Cin, c->Cout, c->KF, c->KT)) tc_total += al(tc_halves(c)) + (comb96(c) ? al(comb_weight_halves(96)) : 0);
        if (tc_total) AID_CUDA_CHECK(cudaMalloc(&n.dweights_tc, tc_total * sizeof(__half)));
        AID_CUDA_CHECK(cudaMalloc(&n.d_sat, 2 * sizeof(unsigned long long)));
        AID_CUDA_CHECK(cudaMemset(n.d_sat, 0, 2 * sizeof(unsigned long long)));
        size_t toff = 0;
        for (auto* c : convs) {
            if (!conv_tc_supported(c->Cin, c->Cout, c->KF, c->KT)) continue;
            Weight& w = n.weights[c->widx];
            AID_CUDA_CHECK(cudaMemcpy(stage, w.host.data(), w.numel() * sizeof(float), cudaMemcpyHostToDevice));
            c->wtc = n.dweights_tc + toff;
            if (n.cfg.conv_mode == 2) launch_pack_weight_tc2(stage, c->wtc, c->Cout, c->Cin, c->KF, c->KT, 0, n.d_sat + 1);
            else launch_pack_weight_tc(stage, c->wtc, c->Cout, c->Cin, c->KF, c->KT, parts, 0);
            AID_CUDA_CHECK(cudaGetLastError());
            AID_CUDA_CHECK(cudaDeviceSynchronize());
            toff += al(tc_halves(c));
            if (comb96(c)) {
                c->wcomb = n.dweights_tc + toff;
                launch_pack_weight_comb(stage, c->wcomb, 96, 0);
                AID_CUDA_CHECK(cudaGetLastError());
                AID_CUDA_CHECK(cudaDeviceSynchronize());
                toff += al(comb_weight_halves(96));
            }
        }
    }
    AID_CUDA_CHECK(cudaFree(stage));
    if (n.d_sat) AID_CUDA_CHECK(cudaMemcpy(&n.weight_sat, n.d_sat + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (auto* nr : norms) {
        Weight& w = n.weights[nr->widx];
        nr->gamma = n.dweights + off;
        AID_CUDA_CHECK(cudaMemcpy(nr->gamma, w.host.data(), w.numel() * sizeof(float), cudaMemcpyHostToDevice));
        off += al(w.numel());
    }
    for (int i = 0; i < 7; ++i) {
        Weight& w = n.weights[n.emb_idx[i]];
        n.d_emb[i] = n.dweights + off;
        AID_CUDA_CHECK(cudaMemcpy(n.d_emb[i], w.host.data(), w.numel() * sizeof(float), cudaMemcpyHostToDevice));
        off += al(w.numel());
    }
    n.d_modW = n.dweights + off; off += al((size_t)n.total_mod * 256);
    n.d_modB = n.dweights + off; off += al((size_t)n.total_mod);
    for (LinRef* l : n.lin_all) {
        AID_CUDA_CHECK(cudaMemcpy(n.d_modW + (size_t)l->off * 256, n.weights[l->w].host.data(), (size_t)l->N * 256 * sizeof(float), cudaMemcpyHostToDevice));
        AID_CUDA_CHECK(cudaMemcpy(n.d_modB + l->off, n.weights[l->b].host.data(), (size_t)l->N * sizeof(float), cudaMemcpyHostToDevice));
    }
    for (auto& w : n.weights) { std::vector<float>().swap(w.host); }
    upload_tables(n);
    n.finalized = true;
}

// CQT tables + FFT twiddles live in one device allocation made at create time (no weights needed).
static void upload_tables(Net& n) {
    if (n.d_tables) return;
    DeviceGuard guard(n.device);
    const CqtPlanHost& p = n.plan;
    const int L = p.L, K = p.K;
    // power-of-two FFT engine size: L itself, or (Bluestein) the next power of two >= 2L
    const bool pow2 = (L & (L - 1)) == 0;
    int M = 1; while (M < (pow2 ? L : 2 * L)) M <<= 1;
    std::vector<float2> tw(M);
    for (int k = 0; k < M; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)M;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<float2> chirp, bfilt;
    if (!pow2) {
        // w[n] = exp(-i*pi*n^2/L) with n^2 reduced mod 2L exactly; filter b[m] = conj(w[|m|]) wrapped on the circle of size M
        chirp.resize(L);
        std::vector<std::complex<double>> wd(L), b(M, 0.0);
        for (long long i = 0; i < L; ++i) {
            const double a = -M_PI * (double)((i * i) % (2LL * L)) / (double)L;
            wd[i] = std::complex<double>(std::cos(a), std::sin(a));
            chirp[i] = make_float2((float)wd[i].real(), (float)wd[i].imag());
        }
        b[0] = std::conj(wd[0]);
        for (int i = 1; i < L; ++i) { b[i] = std::conj(wd[i]); b[M - i] = std::conj(wd[i]); }
        // iterative radix-2 FFT in double on the host (once per handle)
        int lgm = 0; while ((1 << lgm) < M) ++lgm;
        for (int i = 0; i < M; ++i) {
            int r = 0; for (int q = 0; q < lgm; ++q) if (i & (1 << q)) r |= 1 << (lgm - 1 - q);
            if (r > i) std::swap(b[i], b[r]);
        }
        for (int len = 2; len <= M; len <<= 1) {
            const double ang = -2.0 * M_PI / len;
            const std::complex<double> wl(std::cos(ang), std::sin(ang));
            for (int i = 0; i < M; i += len) {
                std::complex<double> w(1.0, 0.0);
                for (int j = 0; j < len / 2; ++j) {
                    const std::complex<double> u = b[i + j], v = b[i + j + len / 2] * w;
                    b[i + j] = u + v; b[i + j + len / 2] = u - v;
                    w *= wl;
                }
            }
        }
        bfilt.resize(M);
        for (int i = 0; i < M; ++i) bfilt[i] = make_float2((float)b[i].real(), (float)b[i].imag());
    }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t bytes = al(M * sizeof(float2)) + al(chirp.size() * sizeof(float2)) + al(bfilt.size() * sizeof(float2)) + 3 * al(K * sizeof(int)) +
                   2 * al(p.win.size() * sizeof(float)) + 2 * al((L / 2 + 1) * sizeof(int)) + al(L * sizeof(float));
    AID_CUDA_CHECK(cudaMalloc(&n.d_tables, bytes));
    char* d = (char*)n.d_tables;
    auto put = [&](const void* src, size_t nb) { void* dst = d; AID_CUDA_CHECK(cudaMemcpy(dst, src, nb, cudaMemcpyHostToDevice)); d += al(nb); return dst; };
    n.fft.tw = (const float2*)put(tw.data(), M * sizeof(float2));
    if (!pow2) {
        n.fft.chirp = (const float2*)put(chirp.data(), chirp.size() * sizeof(float2));
        n.fft.bfilt = (const float2*)put(bfilt.data(), bfilt.size() * sizeof(float2));
    }
    CqtTables& t = n.tabs;
    t.L = L; t.K = K; t.bins = p.bins; t.nocts = p.nocts;
    t.centre = (const int*)put(p.centre.data(), K * sizeof(int));
    t.Lg = (const int*)put(p.Lg.data(), K * sizeof(int));
    t.woff = (const int*)put(p.woff.data(), K * sizeof(int));
    t.win = (const float*)put(p.win.data(), p.win.size() * sizeof(float));
    t.dual = (const float*)put(p.dual.data(), p.dual.size() * sizeof(float));
    t.klo = (const int*)put(p.klo.data(), (L / 2 + 1) * sizeof(int));
    t.khi = (const int*)put(p.khi.data(), (L / 2 + 1) * sizeof(int));
    t.hhpf = (const float*)put(p.hhpf.data(), L * sizeof(float));
}

static void fill_host_tables(Net& n) {
    const CqtPlanHost& p = n.plan;
    {
        const int L = p.L;
        const bool pow2 = (L & (L - 1)) == 0;
        int M = 1; while (M < (pow2 ? L : 2 * L)) M <<= 1;
        int lg = 0; while ((1 << lg) < M) ++lg;
        n.fft.L = L; n.fft.M = M; n.fft.lgM = lg; n.fft.N1 = 1 << ((lg + 1) / 2); n.fft.N2 = M / n.fft.N1;
    }
    CqtTables& t = n.tabs;
    t.L = p.L; t.K = p.K; t.bins = p.bins; t.nocts = p.nocts;
    long long yo = 0;
    for (int o = 0; o < p.nocts; ++o) { t.M[o] = p.size_per_oct[o]; t.yoff[o] = yo; yo += (long long)p.bins * t.M[o]; }
    t.ytotal = yo;
}

// ---- forward ------------------------------------------------------------------------------------------
struct Ctx {
    Net* n = nullptr; Arena ar; cudaStream_t s = nullptr; int B = 0, nsig = 1;
    float* mod = nullptr; double* stats_base = nullptr; int slot = 0;
    Tape* tape = nullptr;   // non-null: taped forward (nothing the backward needs is released or overwritten)
    bool dry() const { return ar.dry; }
    float* allocf(long long nfl) { return (float*)ar.alloc((size_t)nfl * sizeof(float)); }
    void release(void* p) { ar.release(p); }
    double* new_slot() { double* r = stats_base ? stats_base + (size_t)slot * B * 16 : nullptr; ++slot; return dry() ? reinterpret_cast<double*>(0x10) : r; }
    long long modstride() const { return nsig > 1 ? n->total_mod : 0; }
};

#define RUN(call) do { if (!c.dry()) { call; } } while (0)

// fraction of the KF row taps of a dilated convolution that fall inside the plane (the kernels skip the others): the profile
// records count the multiply-adds that are actually executed, not the nominal 2*Cin*Cout*KF*KT per pixel
static double valid_tap_fraction(int F, int KF, int dil) {
    double s = 0;
    for (int kf = 0; kf < KF; ++kf) s += std::max(0, F - std::abs(kf - KF / 2) * dil);
    return s / ((double)KF * F);
}

static void conv(Ctx& c, const TV& a, const ConvW& w, int dil, const TV& out, ConvEpilogue ep) {
    if (a.C != w.Cin || out.C != w.Cout || a.F != out.F || a.T != out.T || a.B != out.B)
        throw std::runtime_error("conv: shape mismatch");
    if (ep.stats && (w.Cout % 8 != 0)) throw std::runtime_error("conv: statistics need Cout % 8 == 0");
    if (c.dry()) return;
    Net& n = *c.n;
    Net::ProfRec rec{};
    if (n.prof) {
        auto get = [&]() { cudaEvent_t e; if (n.prof_pool.empty()) { AID_CUDA_CHECK(cudaEventCreate(&e)); } else { e = n.prof_pool.back(); n.prof_pool.pop_back(); } return e; };
        rec.e0 = get(); rec.e1 = get();
        rec.kind = (w.KF == 5 && w.Cin > 2) ? 0 : 1;   // 0 = dilated 5x3 residual layers (K1), 1 = everything else
        const double px = (double)a.B * a.F * a.T;
        rec.flops = 2.0 * w.Cin * w.Cout * w.KF * w.KT * px * valid_tap_fraction(a.F, w.KF, dil);
        rec.bytes = 4.0 * (px * (w.Cin + w.Cout + (ep.R.p ? w.Cout : 0) + (ep.R2.p ? w.Cout : 0)) + (double)w.Cin * w.Cout * w.KF * w.KT);
        AID_CUDA_CHECK(cudaEventRecord(rec.e0, c.s));
    }
    if (!launch_conv_thin(a, w.wp, w.KF, w.KT, dil, out, ep, c.s)) launch_conv_simt(a, w.wp, w.KF, w.KT, dil, out, ep, c.s);
    if (n.prof) { AID_CUDA_CHECK(cudaEventRecord(rec.e1, c.s)); n.prof_recs.push_back(rec); }
}

static void conv_tc(Ctx& c, const __half* a_hi, const __half* a_lo, int PF, const ConvW& w, int dil, const TV& out, ConvEpilogue ep) {
    if (out.C != w.Cout) throw std::runtime_error("conv_tc: shape mismatch");
    if (c.dry()) return;
    Net& n = *c.n;
    Net::ProfRec rec{};
    if (n.prof) {
        auto get = [&]() { cudaEvent_t e; if (n.prof_pool.empty()) { AID_CUDA_CHECK(cudaEventCreate(&e)); } else { e = n.prof_pool.back(); n.prof_pool.pop_back(); } return e; };
        rec.e0 = get(); rec.e1 = get(); rec.kind = 0;
        const double px = (double)out.B * out.F * out.T;
        rec.kind = (w.KF == 5) ? 0 : 1;
        rec.flops = 2.0 * w.Cin * w.Cout * w.KF * w.KT * px * valid_tap_fraction(out.F, w.KF, dil);
        rec.bytes = 4.0 * (px * (w.Cin + w.Cout + (ep.R.p ? w.Cout : 0)) + (double)w.Cin * w.Cout * w.KF * w.KT);
        AID_CUDA_CHECK(cudaEventRecord(rec.e0, c.s));
    }
    if (n.cfg.conv_mode == 2) launch_conv_tc2(a_hi, PF, w.wtc, out.B, w.Cin, out.F, out.T, w.KF, w.KT, dil, out, ep, n.num_sms, c.s);
    else launch_conv_tc(a_hi, a_lo, PF, w.wtc, out.B, w.Cin, out.F, out.T, w.KF, w.KT, dil, out, ep, n.num_sms, c.s);
    if (n.prof) { AID_CUDA_CHECK(cudaEventRecord(rec.e1, c.s)); n.prof_recs.push_back(rec); }
}

// conv_mode 2, 64 / 96-channel blocks: one fused kernel per dilated layer (conv_comb.cu); recorded as kind 2 (its time includes the
// normalisation / GELU / operand conversion of the layer; bytes = the layer's algorithmic 4 B read + 4 B written per element)
static void conv_comb_layer(Ctx& c, const TV& x, const double* stats_in, long long n_grp, const float* gamma, const float* affine, long long abstride,
                            const ConvW& w, int dil, const TV& out, const ConvEpilogue& ep) {
    if (c.dry()) return;
    Net& n = *c.n;
    Net::ProfRec rec{};
    if (n.prof) {
        auto get = [&]() { cudaEvent_t e; if (n.prof_pool.empty()) { AID_CUDA_CHECK(cudaEventCreate(&e)); } else { e = n.prof_pool.back(); n.prof_pool.pop_back(); } return e; };
        rec.e0 = get(); rec.e1 = get(); rec.kind = 2;
        const double px = (double)out.B * out.F * out.T;
        rec.flops = 2.0 * w.Cin * w.Cout * w.KF * w.KT * px * valid_tap_fraction(out.F, w.KF, dil);
        rec.bytes = 4.0 * px * (w.Cin + w.Cout) + 2.0 * (double)w.Cin * w.Cout * w.KF * w.KT;
        AID_CUDA_CHECK(cudaEventRecord(rec.e0, c.s));
    }
    launch_conv_comb(x, stats_in, n_grp, gamma, affine, abstride, w.wcomb ? w.wcomb : w.wtc, dil, out, ep, n.num_sms, c.s);
    if (n.prof) { AID_CUDA_CHECK(cudaEventRecord(rec.e1, c.s)); n.prof_recs.push_back(rec); }
}

// conv_mode 2, init block of an encoder level with 64 / 96 / 128 channels (conv_init.cu): recorded as kind 1 with the block's algorithmic
// traffic (8 B read per pixel, 4 N B written)
static void init_block_fused(Ctx& c, const ResBlk& k, const TV& in, const TV& out) {
    float* scratch = c.allocf((long long)init_block_scratch_floats(in.B, k.N));
    if (!c.dry()) {
        Net& n = *c.n;
        Net::ProfRec rec{};
        if (n.prof) {
            auto get = [&]() { cudaEvent_t e; if (n.prof_pool.empty()) { AID_CUDA_CHECK(cudaEventCreate(&e)); } else { e = n.prof_pool.back(); n.prof_pool.pop_back(); } return e; };
            rec.e0 = get(); rec.e1 = get(); rec.kind = 1;
            const double px = (double)in.B * in.F * in.T;
            rec.flops = 2.0 * px * k.N * (k.N + 4.0);
            rec.bytes = 4.0 * px * (2.0 + k.N);
            AID_CUDA_CHECK(cudaEventRecord(rec.e0, c.s));
        }
        launch_init_block(in, k.proj_in.wp, k.res_conv.wp, k.H[0].wtc, k.norm[0].gamma, c.mod + k.affine[0].off, c.modstride(),
                          c.mod + k.gate[0].off, c.modstride(), out, out.stats, scratch, n.num_sms, c.s);
        if (n.prof) { AID_CUDA_CHECK(cudaEventRecord(rec.e1, c.s)); n.prof_recs.push_back(rec); }
    }
    c.release(scratch);
}

// conv_mode 2, out blocks (N -> 2 after one gated 1x1 layer; out_block.cu): recorded as kind 1 with the block's algorithmic traffic
static void out_block_fused(Ctx& c, const ResBlk& k, const TV& in, const TV& out, const TV& accum) {
    float* scratch = c.allocf((long long)out_block_scratch_floats(in.B, k.N));
    if (!c.dry()) {
        Net& n = *c.n;
        Net::ProfRec rec{};
        if (n.prof) {
            auto get = [&]() { cudaEvent_t e; if (n.prof_pool.empty()) { AID_CUDA_CHECK(cudaEventCreate(&e)); } else { e = n.prof_pool.back(); n.prof_pool.pop_back(); } return e; };
            rec.e0 = get(); rec.e1 = get(); rec.kind = 1;
            const double px = (double)in.B * in.F * in.T;
            rec.flops = 2.0 * px * k.N * 4.0;
            rec.bytes = 4.0 * px * (k.N + 2.0 + (accum.p ? 2.0 : 0.0));
            AID_CUDA_CHECK(cudaEventRecord(rec.e0, c.s));
        }
        launch_out_block(in, in.stats, (long long)(k.N / 8) * in.F * in.T, k.norm[0].gamma, c.mod + k.affine[0].off, c.modstride(),
                         c.mod + k.gate[0].off, c.modstride(), k.H[0].wp, k.proj_out.wp, k.res_conv.wp, out, accum, scratch, c.s);
        if (n.prof) { AID_CUDA_CHECK(cudaEventRecord(rec.e1, c.s)); n.prof_recs.push_back(rec); }
    }
    c.release(scratch);
}

// unet.py:452-493.  `accum` (decoder out blocks, unet.py:817): out = (accum + block(x)) / sqrt(2), may alias out.
// `bt` (taped forward, input-gradient path): every intermediate the backward needs gets its own buffer and is recorded.
// `up_src` (conv_mode 2, plain forward): the first up_src->C channels of `in` were NOT written; they are the 2x time-upsampling of *up_src,
// which the operand conversion computes on the fly (launch_to_planar_tc2_up) -- the block must consume `in` only through that operand.
static void resblock(Ctx& c, const ResBlk& k, TV in, TV out, const TV* accum = nullptr, BlkTape* bt = nullptr, const TV* up_src = nullptr) {
    const int B = in.B, F = in.F, T = in.T, N = k.N;
    const bool tp = bt != nullptr;
    if (in.C != k.dim || out.C != k.dim_out) throw std::runtime_error("resblock: channel mismatch");
    {   // out blocks (N -> 2 after one gated 1x1 layer): the layer collapses into the projection, one pass over x (AID_OUT_FUSED=0: four launches)
        static const bool env_outb = !(getenv("AID_OUT_FUSED") && atoi(getenv("AID_OUT_FUSED")) == 0);
        if (env_outb && c.n->fuse_out && c.n->cfg.conv_mode == 2 && !tp && !c.n->count_sat && k.after && k.dim_out == 2 && k.k1x1 && k.nd == 1 && !k.attn &&
            k.dim == N && in.stats && !out.stats && out_block_supported(in, out, accum ? *accum : TV())) {
            out_block_fused(c, k, in, out, accum ? *accum : TV());
            return;
        }
    }
    {   // encoder init blocks (2 CQT channels -> N, one gated 1x1 layer): a single fused kernel in conv_mode 2 (AID_INIT_FUSED=0: the five un-fused launches)
        static const bool env_init = !(getenv("AID_INIT_FUSED") && atoi(getenv("AID_INIT_FUSED")) == 0);
        if (env_init && c.n->fuse_init && c.n->cfg.conv_mode == 2 && !tp && !c.n->count_sat && !accum && k.k1x1 && k.dim == 2 && k.nd == 1 && !k.attn && !k.after &&
            k.dim_out == N && k.H[0].wtc && init_block_supported(N, T)) {
            init_block_fused(c, k, in, out);
            return;
        }
    }
    const long long plane = (long long)B * N * F * T;
    const long long n_grp = (long long)(N / 8) * F * T;
    float* xbuf = c.allocf(plane);
    // operand buffer: fp32 [B,N,F,T] for the CUDA-core path, or split-fp16 planar hi|lo [B][N/8][F][T+2][8] each (tcgen05 path)
    // (hi | lo arrays, each with PF zero rows above and below every plane -- PF depends on the layer's dilation)
    auto planar_halves = [&](int ch, int pf) { return (long long)B * ch * (F + 2 * pf) * (T + 2); };
    const int cmode = c.n->cfg.conv_mode;
    const int parts = c.n->tc_parts();
    // fp32 elements of an operand buffer for `ch` channels: conv_mode 1 = (hi | lo) planar arrays, conv_mode 2 = channels-last fp16
    auto operand_floats = [&](int ch, int pf, int Fd) {
        if (cmode == 2) return (long long)((tc2_act_halves(B, ch, Fd, T, pf) + 1) / 2);
        return ((long long)B * ch * (Fd + 2 * pf) * (T + 2) * parts + 1) / 2;
    };
    unsigned long long* sat = c.n->count_sat ? c.n->d_sat : nullptr;
    auto to_operand = [&](const TV& v, int pf, __half* hi, __half* lo) {
        if (cmode == 2) launch_to_planar_tc2(v, pf, hi, c.s, sat); else launch_to_planar_tc(v, pf, hi, lo, c.s);
    };
    const int pf_max = tc_pad_rows(T, k.k1x1 ? 1 : 5, k.k1x1 ? 1 : (1 << std::max(0, k.nd - 1)));
    float* abuf = c.allocf(std::max(plane, operand_floats(N, pf_max, F)));
    __half* a_hi = reinterpret_cast<__half*>(abuf);
    TV x = make_tv(xbuf, B, N, F, T), a = make_tv(abuf, B, N, F, T);
    auto fresh = [&]() { return tp ? make_tv(c.allocf(plane), B, N, F, T) : x; };   // taped: layer outputs never overwrite each other
    if (tp) { bt->in = in; bt->xs.clear(); }
    // tcgen05 path: the block input is converted once to the fp16 operand layout and shared by proj_in and res_conv
    __half* pin_hi = nullptr; __half* pin_lo = nullptr; float* pin_buf = nullptr;
    const int pf1 = tc_pad_rows(T, 1, 1);
    if (k.proj_in.wtc || k.res_conv.wtc) {
        pin_buf = c.allocf(operand_floats(k.dim, pf1, F));
        pin_hi = reinterpret_cast<__half*>(pin_buf);
        pin_lo = parts == 2 ? pin_hi + planar_halves(k.dim, pf1) : nullptr;
        if (up_src) RUN(launch_to_planar_tc2_up(*up_src, slice_c(in, up_src->C, in.C - up_src->C), pf1, pin_hi, c.s));
        else RUN(to_operand(in, pf1, pin_hi, pin_lo));
    }
    if (up_src && !(cmode == 2 && !tp && k.dim != N && k.dim != k.dim_out && k.proj_in.wtc && k.res_conv.wtc && !k.after))
        throw std::runtime_error("resblock: deferred upsampling needs a block that reads its input through the fp16 operand only");
    TV cur;
    if (k.dim != N) {
        TV xo = fresh();
        ConvEpilogue ep; ep.stats = xo.stats = c.new_slot();
        if (k.proj_in.wtc) conv_tc(c, pin_hi, pin_lo, pf1, k.proj_in, 1, xo, ep);
        else conv(c, in, k.proj_in, 1, xo, ep);
        cur = xo;
    } else {
        cur = in;
        if (!cur.stats) { cur.stats = c.new_slot(); RUN(launch_group_stats(cur, cur.stats, c.s)); }
    }
    if (tp) bt->x0 = cur;
    if (k.attn) {
        const int heads = k.a_in.Cout;
        TV h = make_tv(c.allocf((long long)B * heads * F * T), B, heads, F, T);
        // plain forward: the normalisation is folded into the projection's weights (conv_thin.cu); the taped forward keeps the
        // normalised tensor, which its backward reads
        static const bool env_fold = !(getenv("AID_ATT_FOLD") && atoi(getenv("AID_ATT_FOLD")) == 0);
        bool folded = false;
        if (!tp && env_fold && !c.dry())
            folded = launch_conv_thin_out_normed(cur, cur.stats, n_grp, k.norm2.gamma, c.mod + k.affine2.off, c.modstride(), k.a_in.wp, h, c.s);
        if (!folded) {     // (both are no-ops in the planner's dry run)
            RUN(launch_gn_act(cur, cur.stats, n_grp, k.norm2.gamma, c.mod + k.affine2.off, c.modstride(), false, a, c.s));
            conv(c, a, k.a_in, 1, h, ConvEpilogue());
        }
        TV hflat = make_tv(h.p, B, heads * F, 1, T);
        TV qk = make_tv(c.allocf((long long)B * 2 * heads * F * T), B, 2 * heads * F, 1, T);
        if (k.qk.wtc) {
            const long long hh = (long long)B * heads * F * (1 + 2 * pf1) * (T + 2);
            float* hp = c.allocf(operand_floats(heads * F, pf1, 1));
            __half* h_hi = reinterpret_cast<__half*>(hp);
            __half* h_lo = parts == 2 ? h_hi + hh : nullptr;
            RUN(to_operand(hflat, pf1, h_hi, h_lo));
            conv_tc(c, h_hi, h_lo, pf1, k.qk, 1, qk, ConvEpilogue());
            c.release(hp);
        } else {
            conv(c, hflat, k.qk, 1, qk, ConvEpilogue());
        }
        TV o = make_tv(c.allocf((long long)B * heads * F * T), B, heads, F, T);
        static const bool env_att_simt = getenv("AID_ATT_SIMT") && atoi(getenv("AID_ATT_SIMT")) != 0;
        if (cmode == 2 && !env_att_simt && attention_tc_supported(F, T)) RUN(launch_attention_tc(h, qk.p, o, c.s));
        else RUN(launch_attention(h, qk.p, o, c.s));
        ConvEpilogue ep;
        ep.gate = c.mod + k.gate2.off; ep.gate_bstride = c.modstride();
        ep.R = cur; ep.alpha = kInvSqrt2; ep.stats = c.new_slot();
        TV xo = fresh();
        xo.stats = ep.stats;
        conv(c, o, k.a_out, 1, xo, ep);
        if (tp) { bt->h = h; bt->qk = qk; } else { c.release(h.p); c.release(qk.p); }
        c.release(o.p);
        cur = xo;
    }
    if (tp) bt->x1 = cur;
    // conv_mode 2, dilated blocks: between the layers the residual stream lives channels-last ([B][F][T][N] fp32), so that a
    // unit's 128 x N output tile and residual tile are contiguous in HBM (the NCHW planes give 512-byte fragments 1 MB apart,
    // which the DRAM serves at a fraction of its streaming rate).  Layer 0 reads NCHW, the last layer writes NCHW.
    static const bool env_cl = getenv("AID_TC2_CL") && atoi(getenv("AID_TC2_CL")) != 0;   // measured slower than NCHW so far: off by default
    bool use_cl = env_cl && cmode == 2 && !k.k1x1 && k.nd >= 2 && !tp;
    for (auto& h : k.H) use_cl = use_cl && h.wtc != nullptr;
    float* xcl = use_cl ? c.allocf(plane) : nullptr;
    TV xc = make_tv_cl(xcl, B, N, F, T);
    bool cur_is_cl = false;
    for (int i = 0; i < k.nd; ++i) {
        ConvEpilogue ep;
        ep.gate = c.mod + k.gate[i].off; ep.gate_bstride = c.modstride();
        ep.R = cur; ep.alpha = kInvSqrt2;
        ep.stats = (i + 1 < k.nd) ? c.new_slot() : nullptr;
        const double* cur_stats = cur.stats;
        if (tp) { bt->xs.push_back(cur); x = fresh(); }
        if (k.H[i].wtc) {
            const int dil = k.k1x1 ? 1 : (1 << i);
            const int pf = tc_pad_rows(T, k.H[i].KF, dil);
            __half* a_lo = parts == 2 ? a_hi + planar_halves(N, pf) : nullptr;
            static const bool env_comb = !(getenv("AID_COMB") && atoi(getenv("AID_COMB")) == 0);
            // 96 channels (AID_COMB96=0 turns it off): the fused kernel streams its 270 KB of weights and is bound by shared-memory
            // bandwidth like conv_tc2; in the network at batch 32 it measured +1.4 % over pass + conv_tc2 (cta_group::2), and it differs
            // from that path in fp32 accumulation order (2e-6 per layer)
            static const bool env_comb96 = !(getenv("AID_COMB96") && atoi(getenv("AID_COMB96")) == 0);
            if (cmode == 2 && env_comb && c.n->fuse_comb && !tp && !use_cl && !k.k1x1 && !sat && conv_comb_supported(N, F, T, dil) && (N != 96 || (env_comb96 && k.H[i].wcomb)) &&
                conv_comb_worthwhile(B, T, dil, c.n->num_sms)) {
                // fused layer (conv_comb.cu): normalisation, modulation, GELU and the operand conversion happen inside the convolution;
                // the t-tile halos forbid an in-place update, so the layers alternate between the block's two buffers
                TV o = (cur.p == x.p) ? a : x;
                o.stats = ep.stats;
                conv_comb_layer(c, cur, cur_stats, n_grp, k.norm[i].gamma, c.mod + k.affine[i].off, c.modstride(), k.H[i], dil, o, ep);
                cur = o;
                continue;
            }
            // (the fused layers are the LAST ones of a block -- a comb gets more parallel with the dilation -- so the operand buffer is never
            // needed again once one of them has used it as its output plane)
            if (cur.p == abuf) throw std::runtime_error("resblock: un-fused layer after a fused one");
            if (cmode == 2) {
                if (cur_is_cl) RUN(launch_gn_act_tc2_cl(xcl, B, N, F, T, cur_stats, n_grp, k.norm[i].gamma, c.mod + k.affine[i].off, c.modstride(), true, pf, a_hi, c.s));
                else RUN(launch_gn_act_tc2(cur, cur_stats, n_grp, k.norm[i].gamma, c.mod + k.affine[i].off, c.modstride(), true, pf, a_hi, c.s, sat));
                const bool out_is_cl = use_cl && i + 1 < k.nd;
                if (cur_is_cl) ep.R = xc;
                ep.R_cl = cur_is_cl; ep.out_cl = out_is_cl;
                x.stats = ep.stats;
                conv_tc(c, a_hi, nullptr, pf, k.H[i], dil, out_is_cl ? xc : x, ep);
                cur_is_cl = out_is_cl;
            } else {
                RUN(launch_gn_act_tc(cur, cur_stats, n_grp, k.norm[i].gamma, c.mod + k.affine[i].off, c.modstride(), true, pf, a_hi, a_lo, c.s));
                x.stats = ep.stats;
                conv_tc(c, a_hi, a_lo, pf, k.H[i], dil, x, ep);
            }
        } else {
            RUN(launch_gn_act(cur, cur_stats, n_grp, k.norm[i].gamma, c.mod + k.affine[i].off, c.modstride(), true, a, c.s));
            x.stats = ep.stats;
            conv(c, a, k.H[i], k.k1x1 ? 1 : (1 << i), x, ep);
        }
        cur = x;
    }
    if (xcl) c.release(xcl);
    if (k.after && N != k.dim_out) {
        TV t = make_tv(c.allocf((long long)B * k.dim_out * F * T), B, k.dim_out, F, T);
        conv(c, cur, k.proj_out, 1, t, ConvEpilogue());
        ConvEpilogue ep; ep.R = t; ep.stats = out.stats;
        if (accum) { ep.alpha = 0.5f; ep.beta = kInvSqrt2; ep.R2 = *accum; }
        else ep.alpha = kInvSqrt2;
        conv(c, in, k.res_conv, 1, out, ep);
        c.release(t.p);
    } else {
        if (accum) throw std::runtime_error("resblock: accum only supported for out blocks");
        if (k.dim != k.dim_out) {
            ConvEpilogue ep; ep.R = cur; ep.alpha = kInvSqrt2; ep.stats = out.stats;
            if (k.res_conv.wtc) conv_tc(c, pin_hi, pin_lo, pf1, k.res_conv, 1, out, ep);
            else conv(c, in, k.res_conv, 1, out, ep);
        } else {
            RUN(launch_combine(cur, in, kInvSqrt2, kInvSqrt2, out, out.stats, c.s));
        }
    }
    c.release(abuf); c.release(xbuf);
    if (pin_buf) c.release(pin_buf);
}

// debug probes (parity tests localise an error to one block): contiguous copy of a named intermediate
static void probe(Ctx& c, const std::string& name, const TV& v) {
    if (c.dry()) return;
    auto it = c.n->probes.find(name);
    if (it == c.n->probes.end() || !it->second) return;
    launch_combine(v, TV(), 1.f, 0.f, make_tv(it->second, v.B, v.C, v.F, v.T), nullptr, c.s);
}

// unet.py:730-845
static void forward(Ctx& c, const float* x, const float* c_noise, float* out, float in_scale, float out_scale, float skip_scale,
                    const float* dscal = nullptr) {
    Net& n = *c.n;
    FftPlan fftd = n.fft; fftd.dscal = dscal;   // in / out / skip scales from device memory (aid_unet_forward_ds)
    const aid_config& cf = n.cfg;
    const int B = c.B, L = cf.audio_len, no = cf.num_octs, bins = cf.bins_per_oct;
    c.slot = 0;
    Tape* tp = c.tape;
    auto keep = [&](void* p) { if (!tp) c.release(p); };   // released in the plain forward, kept for the backward in the taped one
    if (tp) {
        tp->init.assign(no, BlkTape()); tp->main.assign(no, BlkTape()); tp->ups_main.assign(no, BlkTape()); tp->ups_out.assign(no, BlkTape());
        tp->mid_main = BlkTape(); tp->mid_out = BlkTape();
    }
    c.mod = c.allocf((long long)B * n.total_mod);   // sized for per-clip sigma so the plan does not depend on n_sigma
    float* emb = c.allocf((long long)B * 256);
    RUN(launch_embedding(c_noise, c.nsig, n.d_emb[0], n.d_emb[1], n.d_emb[2], n.d_emb[3], n.d_emb[4], n.d_emb[5], n.d_emb[6], emb, c.s));
    RUN(launch_mod_vectors(emb, c.nsig, n.d_modW, n.d_modB, n.total_mod, c.mod, c.s));
    const size_t stat_bytes = (size_t)std::max(1, n.n_stat_slots) * B * 16 * sizeof(double);
    c.stats_base = (double*)c.ar.alloc(stat_bytes);
    RUN(AID_CUDA_CHECK(cudaMemsetAsync(c.stats_base, 0, stat_bytes, c.s)));
    float2* spec = (float2*)c.ar.alloc((size_t)B * L * sizeof(float2));
    float2* tmp = (float2*)c.ar.alloc((size_t)B * n.fft.M * sizeof(float2));
    float2* fscr = n.fft.M != L ? (float2*)c.ar.alloc((size_t)B * n.fft.M * sizeof(float2)) : nullptr;
    RUN(launch_fft_big(fftd, B, false, x, L, in_scale, nullptr, tmp, fscr, spec, nullptr, 0, 0.f, nullptr, 0, 0.f, c.s));

    auto Tof = [&](int lvl) { return n.tabs.M[no - 1 - lvl]; };
    std::vector<TV> cat(no);
    for (int j = 0; j < no; ++j)
        cat[j] = make_tv(c.allocf((long long)B * 2 * cf.Ns[j] * bins * (j + 1) * Tof(j)), B, 2 * cf.Ns[j], bins * (j + 1), Tof(j));

    TV Xcat, pyr, Xmid;
    for (int i = 0; i < no; ++i) {
        const int Ti = Tof(i), Fi = bins * (i + 1);
        const int din = i == 0 ? cf.Ns[0] : cf.Ns[i - 1];
        TV C = make_tv(c.allocf((long long)B * 2 * bins * Ti), B, 2, bins, Ti);
        RUN(launch_cqt_analysis_oct(n.tabs, n.fft, no - 1 - i, spec, C, c.s));
        if (i == 0) { Xcat = make_tv(c.allocf((long long)B * din * Fi * Ti), B, din, Fi, Ti); Xcat.stats = c.new_slot(); }
        TV c2 = slice_f(Xcat, 0, bins); c2.stats = Xcat.stats;
        resblock(c, n.downs[i].init, C, c2, nullptr, tp ? &tp->init[i] : nullptr);
        TV pyr_new;
        if (i < no - 1) {
            pyr_new = make_tv(c.allocf((long long)B * 2 * Fi * (Ti / 2)), B, 2, Fi, Ti / 2);
            RUN(launch_resample_down(C, slice_f(pyr_new, 0, bins), c.s));
            if (i > 0) RUN(launch_resample_down(pyr, slice_f(pyr_new, bins, Fi - bins), c.s));
        } else {
            pyr_new = make_tv(c.allocf((long long)B * 2 * Fi * Ti), B, 2, Fi, Ti);
            RUN(launch_combine(C, TV(), 1.f, 0.f, slice_f(pyr_new, 0, bins), nullptr, c.s));
            if (i > 0) RUN(launch_combine(pyr, TV(), 1.f, 0.f, slice_f(pyr_new, bins, Fi - bins), nullptr, c.s));
        }
        keep(C.p);
        if (i > 0) c.release(pyr.p);
        pyr = pyr_new;
        TV skip = slice_c(cat[i], cf.Ns[i], cf.Ns[i]);
        resblock(c, n.downs[i].main, Xcat, skip, nullptr, tp ? &tp->main[i] : nullptr);
        probe(c, "enc" + std::to_string(i), skip);
        keep(Xcat.p);
        ConvEpilogue ep; ep.alpha = kInvSqrt2;
        if (i < no - 1) {
            TV Xd = make_tv(c.allocf((long long)B * cf.Ns[i] * Fi * (Ti / 2)), B, cf.Ns[i], Fi, Ti / 2);
            RUN(launch_resample_down(skip, Xd, c.s));
            Xcat = make_tv(c.allocf((long long)B * cf.Ns[i] * (Fi + bins) * (Ti / 2)), B, cf.Ns[i], Fi + bins, Ti / 2);
            Xcat.stats = c.new_slot();
            ep.R = Xd; ep.stats = Xcat.stats;
            conv(c, pyr, n.downs[i].pyr, 1, slice_f(Xcat, bins, Fi), ep);
            c.release(Xd.p);
        } else {
            Xmid = make_tv(c.allocf((long long)B * cf.Ns[i] * Fi * Ti), B, cf.Ns[i], Fi, Ti);
            Xmid.stats = c.new_slot();
            ep.R = skip; ep.stats = Xmid.stats;
            conv(c, pyr, n.downs[i].pyr, 1, Xmid, ep);
        }
    }
    c.release(pyr.p);

    const int Fl = bins * no, Tl = Tof(no - 1);
    TV Xm = slice_c(cat[no - 1], 0, cf.Ns[no - 1]); Xm.stats = c.new_slot();
    resblock(c, n.mid_main, Xmid, Xm, nullptr, tp ? &tp->mid_main : nullptr);
    probe(c, "mid", Xm);
    keep(Xmid.p);
    TV Xout = make_tv(c.allocf((long long)B * 2 * Fl * Tl), B, 2, Fl, Tl);
    resblock(c, n.mid_out, Xm, Xout, nullptr, tp ? &tp->mid_out : nullptr);
    float2* Y = (float2*)c.ar.alloc((size_t)B * n.tabs.ytotal * sizeof(float2));
    TV up_src; float* up_hold = nullptr;           // deferred upsampling of the decoder stream (see below)
    for (int i = 0; i < no; ++i) {
        const int j = no - 1 - i, Fj = bins * (j + 1), Tj = Tof(j);
        const int dout = j == 0 ? cf.Ns[0] : cf.Ns[j - 1];
        TV Xdec = make_tv(c.allocf((long long)B * dout * Fj * Tj), B, dout, Fj, Tj); Xdec.stats = c.new_slot();
        resblock(c, n.ups_main[i], cat[j], Xdec, nullptr, tp ? &tp->ups_main[i] : nullptr, up_hold ? &up_src : nullptr);
        if (up_hold) { c.release(up_hold); up_hold = nullptr; }
        probe(c, "dec" + std::to_string(i), Xdec);
        keep(cat[j].p);
        resblock(c, n.ups_out[i], Xdec, Xout, &Xout, tp ? &tp->ups_out[i] : nullptr);
        RUN(launch_cqt_synth_oct(n.tabs, n.fft, i, slice_f(Xout, 0, bins), Y, c.s));
        if (j > 0) {
            TV Xn = slice_c(cat[j - 1], 0, cf.Ns[j - 1]);
            // conv_mode 2: the upsampled decoder stream is consumed only as the fp16 operand of the next level's proj_in / res_conv, so it is
            // produced there, inside the operand conversion, from the half-rate rows (AID_UP_FUSED=0: resample_up + conversion of the fp32 copy)
            static const bool env_up = !(getenv("AID_UP_FUSED") && atoi(getenv("AID_UP_FUSED")) == 0);
            const ResBlk& nb = n.ups_main[i + 1];
            const TV usrc = slice_f(Xdec, bins, Fj - bins);
            if (env_up && n.fuse_up && cf.conv_mode == 2 && !tp && !n.count_sat && nb.proj_in.wtc && nb.res_conv.wtc && !nb.after && nb.dim != nb.N &&
                nb.dim != nb.dim_out && to_planar_tc2_up_supported(usrc, slice_c(cat[j - 1], cf.Ns[j - 1], cat[j - 1].C - cf.Ns[j - 1]))) {
                up_src = usrc; up_hold = Xdec.p;       // Xdec stays alive until the next main block has converted its input
            } else {
                RUN(launch_resample_up(usrc, Xn, c.s));
            }
            TV Xo2 = make_tv(c.allocf((long long)B * 2 * (Fj - bins) * 2 * Tj), B, 2, Fj - bins, 2 * Tj);
            RUN(launch_resample_up(slice_f(Xout, bins, Fj - bins), Xo2, c.s));
            c.release(Xout.p);
            Xout = Xo2;
        }
        if (!up_hold) keep(Xdec.p);
    }
    c.release(Xout.p);
    RUN(launch_cqt_synth_gather(n.tabs, B, Y, spec, c.s));
    RUN(launch_fft_big(fftd, B, true, nullptr, 0, 0.f, spec, tmp, fscr, nullptr, out, L, out_scale / (float)L,
                       (skip_scale != 0.f || dscal) ? x : nullptr, L, skip_scale, c.s));
    if (tp) { tp->mod = c.mod; tp->spec = spec; tp->tmp = tmp; tp->fscr = fscr; tp->Y = Y; }
    if (!c.dry()) AID_CUDA_CHECK(cudaGetLastError());
}

static size_t plan_forward(Net& n, int B, int* slots) {
    Ctx c; c.n = &n; c.B = B; c.nsig = 1; c.ar.reset(nullptr, 0, true);
    forward(c, nullptr, nullptr, nullptr, 1.f, 1.f, 0.f);
    if (slots) *slots = c.slot;
    return c.ar.peak + 4096;
}

// ======================================================================================================
// Input gradient (VJP) of the denoiser: reconstruction guidance, sampler.py:57-113 (torch.autograd.grad(norm, x) through
// edm.py:133-148 and unet.py:730-845).  The taped forward keeps every block input / layer input alive in the workspace; the
// backward below walks the network in reverse.  The data-gradient convolutions are the forward convolution kernels run with
// transposed, tap-mirrored weights: fp32 CUDA cores in conv_mode 0 / 1 and for the thin-channel layers, tcgen05 with fp16 operands
// and per-tensor gradient scaling for the dilated and the wide 1x1 layers in conv_mode 2; everything else is fp32.
// Convention: every backward routine ACCUMULATES into its destination (zero-initialised by the caller), through the
// convolution epilogue's R2 input or the beta argument of the adjoint kernels.

// conv_mode 2: does the data gradient of this convolution run on tcgen05 (fp16 operands, transposed weights wtcT built by ensure_vjp)?
// Depends only on what is known at planning time (the planner's dry run may precede ensure_vjp).
static bool vjp_on_tc(const Net& n, const ConvW& w) {
    static const bool env_tc1 = !(getenv("AID_VJP_TC1X1") && atoi(getenv("AID_VJP_TC1X1")) == 0);
    const bool shape = (w.KF == 5 && w.KT == 3) || (env_tc1 && w.KF == 1 && w.KT == 1);
    return n.cfg.conv_mode == 2 && shape && w.wtc != nullptr && conv_tc_supported(w.Cout, w.Cin, w.KF, w.KT);
}

// transposed weights + CQT adjoint tables, built once per handle on the first VJP call (allocates; documented in the header)
static void ensure_vjp(Net& n) {
    if (n.vjp_ready) return;
    DeviceGuard guard(n.device);
    std::vector<ConvW*> convs;
    auto visit = [&](ResBlk& k) {
        for (ConvW* c : {&k.proj_in, &k.res_conv, &k.proj_out, &k.a_in, &k.a_out, &k.qk}) if (c->widx >= 0) convs.push_back(c);
        for (auto& c : k.H) convs.push_back(&c);
    };
    for (auto& l : n.downs) { visit(l.init); visit(l.main); convs.push_back(&l.pyr); }
    visit(n.mid_main); visit(n.mid_out);
    for (auto& k : n.ups_main) visit(k);
    for (auto& k : n.ups_out) visit(k);
    auto al = [](size_t v) { return (v + 63) & ~(size_t)63; };
    size_t total = 0;
    for (auto* c : convs) total += al((size_t)c->Cout * c->Cin * c->KF * c->KT);
    AID_CUDA_CHECK(cudaMalloc(&n.dweights_T, total * sizeof(float)));
    size_t off = 0;
    for (auto* c : convs) {
        c->wpT = n.dweights_T + off;
        launch_pack_conv_weight_T(c->wp, c->wpT, c->Cout, c->Cin, c->KF, c->KT, nullptr);
        off += al((size_t)c->Cout * c->Cin * c->KF * c->KT);
    }
    AID_CUDA_CHECK(cudaGetLastError());
    if (n.cfg.conv_mode == 2) {
        // the data gradient of the dilated 5x3 layers (95 % of the backward FLOPs) and of the wide 1x1 convolutions runs on tcgen05 too
        size_t tct = 0, stage_n = 0;
        for (auto* c : convs)
            if (vjp_on_tc(n, *c)) {
                tct += al(tc2_weight_halves(c->Cin, c->Cout, c->KF, c->KT)); stage_n = std::max(stage_n, (size_t)c->Cout * c->Cin * c->KF * c->KT);
            }
        if (tct) {
            float* stage = nullptr;
            AID_CUDA_CHECK(cudaMalloc(&n.dweights_tcT, tct * sizeof(__half)));
            AID_CUDA_CHECK(cudaMalloc(&stage, stage_n * sizeof(float)));
            size_t toff = 0;
            for (auto* c : convs) {
                if (!vjp_on_tc(n, *c)) continue;
                launch_transpose_weight_std(c->wp, stage, c->Cout, c->Cin, c->KF * c->KT, nullptr);
                c->wtcT = n.dweights_tcT + toff;
                launch_pack_weight_tc2(stage, c->wtcT, /*Cout'=*/c->Cin, /*Cin'=*/c->Cout, c->KF, c->KT, nullptr, nullptr);
                AID_CUDA_CHECK(cudaDeviceSynchronize());
                toff += al(tc2_weight_halves(c->Cin, c->Cout, c->KF, c->KT));
            }
            AID_CUDA_CHECK(cudaFree(stage));
        }
    }
    const CqtPlanHost& p = n.plan;
    std::vector<float> winM(p.win.size()), dualM(p.dual.size());
    for (int k = 0; k < p.K; ++k) {
        const float M = (float)p.size_per_oct[k / p.bins];
        for (int i = 0; i < p.Lg[k]; ++i) { winM[p.woff[k] + i] = p.win[p.woff[k] + i] / M; dualM[p.woff[k] + i] = p.dual[p.woff[k] + i] * M; }
    }
    AID_CUDA_CHECK(cudaMalloc(&n.d_adj_tables, (winM.size() + dualM.size()) * sizeof(float)));
    AID_CUDA_CHECK(cudaMemcpy(n.d_adj_tables, winM.data(), winM.size() * sizeof(float), cudaMemcpyHostToDevice));
    AID_CUDA_CHECK(cudaMemcpy(n.d_adj_tables + winM.size(), dualM.data(), dualM.size() * sizeof(float), cudaMemcpyHostToDevice));
    n.tabs_adjA = n.tabs; n.tabs_adjA.dual = n.d_adj_tables;                    // gather with the analysis windows / M
    n.tabs_adjS = n.tabs; n.tabs_adjS.win = n.d_adj_tables + winM.size();       // window-and-fold with the dual windows * M
    AID_CUDA_CHECK(cudaDeviceSynchronize());
    n.vjp_ready = true;
}

static TV alloc_tv(Ctx& c, int C, int F, int T, bool zero) {
    TV v = make_tv(c.allocf((long long)c.B * C * F * T), c.B, C, F, T);
    if (zero) RUN(AID_CUDA_CHECK(cudaMemsetAsync(v.p, 0, (size_t)c.B * C * F * T * sizeof(float), c.s)));
    return v;
}

// dst += alpha * conv_transposed(g)      (dst: [B, w.Cin, F, T], g: [B, w.Cout, F, T])
static void conv_T(Ctx& c, const TV& g, const ConvW& w, int dil, const TV& dst, float alpha) {
    if (w.KF == 1 && w.KT == 1 && vjp_on_tc(*c.n, w)) {
        // 1x1 data gradient on tcgen05 (round 2: these were 44 % of the backward on fp32 CUDA cores): the gradient is scaled per tensor by
        // a power of two found on the device (max |g| -> [96, 192)) so that its fp16 image keeps full relative precision; the epilogue
        // undoes the scale, applies alpha and accumulates into dst (out = acc * inv_vec + dst, in place)
        const int Ci = w.Cout, Co = w.Cin;      // channels of g / of dst
        const int pf = tc_pad_rows(g.T, 1, 1);
        float* ab = c.allocf((long long)((tc2_act_halves(c.B, Ci, g.F, g.T, pf) + 1) / 2));
        float* sc = c.allocf(2 + Co);
        unsigned int* amax = reinterpret_cast<unsigned int*>(sc); float* scal = sc + 1; float* inv_vec = sc + 2;
        RUN(AID_CUDA_CHECK(cudaMemsetAsync(amax, 0, sizeof(unsigned int), c.s)));
        RUN(launch_absmax(g, amax, c.s));
        RUN(launch_tc_scale(amax, 1.f, alpha, scal, inv_vec, Co, c.s));
        RUN(launch_gn_act_tc2(g, nullptr, 1, nullptr, scal, 0, false, pf, reinterpret_cast<__half*>(ab), c.s, nullptr));
        ConvEpilogue ep; ep.gate = inv_vec; ep.gate_bstride = 0; ep.alpha = 1.f; ep.R = dst;
        if (!c.dry()) launch_conv_tc2(reinterpret_cast<__half*>(ab), pf, w.wtcT, c.B, Ci, g.F, g.T, 1, 1, 1, dst, ep, c.n->num_sms, c.s);
        c.release(ab); c.release(sc);
        return;
    }
    ConvW wt; wt.Cin = w.Cout; wt.Cout = w.Cin; wt.KF = w.KF; wt.KT = w.KT; wt.wp = w.wpT; wt.widx = w.widx;
    ConvEpilogue ep; ep.alpha = alpha; ep.beta = 1.f; ep.R2 = dst;
    conv(c, g, wt, dil, dst, ep);
}

// unet.py:353-374 backward.  h [B,heads,F,T], qk [B,2*heads*F,1,T], g_o [B,heads,F,T] -> g_h (overwritten) [B,heads,F,T]
static void attention_core_bwd(Ctx& c, const TV& h, const TV& qk, const TV& g_o, const TV& g_h, const TV& gqk) {
    const int B = c.B, heads = h.C, F = h.F, T = h.T, Z = B * heads;
    const long long FT = (long long)F * T, TT = (long long)T * T;
    const float scale = 1.0f / sqrtf((float)F);
    float* P = c.allocf((long long)Z * TT);
    float* gP = c.allocf((long long)Z * TT);
    const float* Qt = qk.p;            // [z][d][t] at z * 2FT
    const float* Kt = qk.p + FT;
    BGemm g{};
    g.batch = Z; g.beta = 0.f;
    // S[t,tk] = scale * sum_d q[t,d] k[tk,d]
    g.A = Qt; g.sAm = 1; g.sAk = T; g.sAz = 2 * FT; g.B = Kt; g.sBk = T; g.sBn = 1; g.sBz = 2 * FT;
    g.C = P; g.sCm = T; g.sCn = 1; g.sCz = TT; g.M = T; g.N = T; g.K = F; g.alpha = scale;
    RUN(launch_bgemm(g, c.s));
    RUN(launch_softmax_rows(P, (long long)Z * T, T, c.s));
    // gP[t,tk] = sum_f g_o[f,t] h[f,tk]
    g.A = g_o.p; g.sAm = 1; g.sAk = T; g.sAz = FT; g.B = h.p; g.sBk = T; g.sBn = 1; g.sBz = FT;
    g.C = gP; g.sCm = T; g.sCn = 1; g.sCz = TT; g.M = T; g.N = T; g.K = F; g.alpha = 1.f;
    RUN(launch_bgemm(g, c.s));
    RUN(launch_softmax_bwd(P, gP, (long long)Z * T, T, scale, c.s));     // gP <- gS (includes the logit scale)
    // g_q^T[d,t] = sum_tk k^T[d,tk] gS[t,tk]
    g.A = Kt; g.sAm = T; g.sAk = 1; g.sAz = 2 * FT; g.B = gP; g.sBk = 1; g.sBn = T; g.sBz = TT;
    g.C = gqk.p; g.sCm = T; g.sCn = 1; g.sCz = 2 * FT; g.M = F; g.N = T; g.K = T; g.alpha = 1.f;
    RUN(launch_bgemm(g, c.s));
    // g_k^T[d,tk] = sum_t q^T[d,t] gS[t,tk]
    g.A = Qt; g.sAm = T; g.sAk = 1; g.sAz = 2 * FT; g.B = gP; g.sBk = T; g.sBn = 1; g.sBz = TT;
    g.C = gqk.p + FT; g.M = F; g.N = T; g.K = T;
    RUN(launch_bgemm(g, c.s));
    // value path: g_h[f,tk] = sum_t g_o[f,t] P[t,tk]
    g.A = g_o.p; g.sAm = T; g.sAk = 1; g.sAz = FT; g.B = P; g.sBk = T; g.sBn = 1; g.sBz = TT;
    g.C = g_h.p; g.sCm = T; g.sCn = 1; g.sCz = FT; g.M = F; g.N = T; g.K = T;
    RUN(launch_bgemm(g, c.s));
    c.release(P); c.release(gP);
}
static void attention_bwd(Ctx& c, const ResBlk& k, const TV& h, const TV& qk, const TV& g_o, const TV& g_h) {
    const int heads = h.C, F = h.F, T = h.T;
    TV gqk = alloc_tv(c, 2 * heads * F, 1, T, false);
    attention_core_bwd(c, h, qk, g_o, g_h, gqk);
    // q|k path: g_h (as [B, heads*F, 1, T]) += Wqk^T g_qk
    conv_T(c, gqk, k.qk, 1, make_tv(g_h.p, c.B, heads * F, 1, T), 1.f);
    c.release(gqk.p);
}

// Backward of resblock(): g_in += d out / d in ^T (a_tail-scaled g_out).  a_tail: the coefficient of the block's own path in its
// output (1/sqrt 2, or 0.5 for the decoder out blocks whose output is (accum + block) / sqrt 2 after the block's own / sqrt 2).
static void resblock_bwd(Ctx& c, const ResBlk& k, const BlkTape& bt, const TV& g_out, float a_tail, const TV& g_in) {
    const int F = g_out.F, T = g_out.T, N = k.N;
    const long long n_grp = (long long)(N / 8) * F * T;
    double* D = (double*)c.ar.alloc((size_t)c.B * 8 * sizeof(double));
    TV g_cur = alloc_tv(c, N, F, T, false);
    if (k.after && N != k.dim_out) {
        conv_T(c, g_out, k.res_conv, 1, g_in, a_tail);
        RUN(AID_CUDA_CHECK(cudaMemsetAsync(g_cur.p, 0, (size_t)c.B * N * F * T * sizeof(float), c.s)));
        conv_T(c, g_out, k.proj_out, 1, g_cur, a_tail);
    } else if (k.dim != k.dim_out) {
        conv_T(c, g_out, k.res_conv, 1, g_in, a_tail);
        RUN(launch_scale_channels(g_out, nullptr, 0, a_tail, g_cur, c.s));
    } else {
        RUN(launch_combine(g_out, g_in, a_tail, 1.f, g_in, nullptr, c.s));
        RUN(launch_scale_channels(g_out, nullptr, 0, a_tail, g_cur, c.s));
    }
    TV tmp = alloc_tv(c, N, F, T, false), ga = alloc_tv(c, N, F, T, false);
    // conv_mode 2: the data gradient of the dilated layers on tcgen05.  The gradient is scaled per tensor by a power of two
    // (max |g| -> [96, 192), found on the device) so that its fp16 image keeps full relative precision, multiplied by the layer's
    // gate per channel in the same pass, and the convolution's epilogue undoes the scale.
    bool tc = c.n->cfg.conv_mode == 2 && !k.k1x1 && k.nd > 0;
    for (auto& hw : k.H) tc = tc && hw.wtc != nullptr && conv_tc_supported(hw.Cout, hw.Cin, hw.KF, hw.KT);   // == ensure_vjp's condition for wtcT
    __half* a16 = nullptr; unsigned int* amax = nullptr; float* scal = nullptr; float* inv_vec = nullptr; float* abuf16 = nullptr;
    if (tc) {
        const int pf_max = tc_pad_rows(T, 5, 1 << std::max(0, k.nd - 1));
        abuf16 = c.allocf((long long)((tc2_act_halves(c.B, N, F, T, pf_max) + 1) / 2));
        a16 = reinterpret_cast<__half*>(abuf16);
        float* sc = c.allocf(2 + N);
        amax = reinterpret_cast<unsigned int*>(sc); scal = sc + 1; inv_vec = sc + 2;
        RUN(AID_CUDA_CHECK(cudaMemsetAsync(amax, 0, sizeof(unsigned int), c.s)));
        RUN(launch_absmax(g_cur, amax, c.s));
    }
    for (int i = k.nd - 1; i >= 0; --i) {
        // x_{i+1} = (x_i + H_i(gelu(GN_i(x_i) (1 + affine_i))) gate_i) / sqrt 2
        const TV& xi = bt.xs[i];
        const int dil = k.k1x1 ? 1 : (1 << i);
        if (tc) {
            const int pf = tc_pad_rows(T, 5, dil);
            RUN(launch_tc_scale(amax, kInvSqrt2, 1.f, scal, inv_vec, N, c.s));
            RUN(launch_gn_act_tc2(g_cur, nullptr, 1, c.mod + k.gate[i].off, scal, c.modstride(), false, pf, a16, c.s, nullptr));
            ConvEpilogue ep; ep.gate = inv_vec; ep.gate_bstride = 0; ep.alpha = 1.f;
            if (!c.dry()) launch_conv_tc2(a16, pf, k.H[i].wtcT, c.B, N, F, T, 5, 3, dil, ga, ep, c.n->num_sms, c.s);
        } else {
            RUN(launch_scale_channels(g_cur, c.mod + k.gate[i].off, c.modstride(), kInvSqrt2, tmp, c.s));
            RUN(AID_CUDA_CHECK(cudaMemsetAsync(ga.p, 0, (size_t)c.B * N * F * T * sizeof(float), c.s)));
            conv_T(c, tmp, k.H[i], dil, ga, 1.f);
        }
        RUN(launch_gn_bwd(ga, xi, xi.stats, n_grp, k.norm[i].gamma, c.mod + k.affine[i].off, c.modstride(), true, D, g_cur, kInvSqrt2, g_cur, c.s,
                          tc && i > 0 ? amax : nullptr));
    }
    if (tc) { c.release(abuf16); c.release(reinterpret_cast<float*>(amax)); }
    if (k.attn) {
        // x1 = (x0 + a_out(attention(a_in(GN2(x0) (1 + affine2)))) gate2) / sqrt 2
        const int heads = k.a_in.Cout;
        RUN(launch_scale_channels(g_cur, c.mod + k.gate2.off, c.modstride(), kInvSqrt2, tmp, c.s));
        TV g_o = alloc_tv(c, heads, F, T, true), g_h = alloc_tv(c, heads, F, T, false);
        conv_T(c, tmp, k.a_out, 1, g_o, 1.f);
        attention_bwd(c, k, bt.h, bt.qk, g_o, g_h);
        RUN(AID_CUDA_CHECK(cudaMemsetAsync(ga.p, 0, (size_t)c.B * N * F * T * sizeof(float), c.s)));
        conv_T(c, g_h, k.a_in, 1, ga, 1.f);
        RUN(launch_gn_bwd(ga, bt.x0, bt.x0.stats, n_grp, k.norm2.gamma, c.mod + k.affine2.off, c.modstride(), false, D, g_cur, kInvSqrt2, g_cur, c.s));
        c.release(g_o.p); c.release(g_h.p);
    }
    if (k.dim != N) conv_T(c, g_cur, k.proj_in, 1, g_in, 1.f);
    else RUN(launch_combine(g_cur, g_in, 1.f, 1.f, g_in, nullptr, c.s));
    c.release(tmp.p); c.release(ga.p); c.release(g_cur.p); c.release(D);
}

// g_x = in_scale * out_scale * J_net^T g_out + skip_scale * g_out, for the forward recorded in the tape
static void backward(Ctx& c, Tape& tp, const float* g_out, float* g_x) {
    Net& n = *c.n;
    const aid_config& cf = n.cfg;
    const int B = c.B, L = cf.audio_len, no = cf.num_octs, bins = cf.bins_per_oct;
    auto Tof = [&](int lvl) { return n.tabs.M[no - 1 - lvl]; };
    // ---- synthesis adjoint: gradient of every octave's output coefficients (unet.py:841, 826-836) ----
    RUN(launch_fft_big(n.fft, B, false, g_out, L, 1.f, nullptr, tp.tmp, tp.fscr, tp.spec, nullptr, 0, 0.f, nullptr, 0, 0.f, c.s));
    RUN(launch_spec_synth_adj(B, L, tp.spec, tp.out_scale / (float)L, c.s));
    std::vector<TV> gTop(no);
    for (int i = 0; i < no; ++i) {
        gTop[i] = alloc_tv(c, 2, bins, n.tabs.M[i], false);
        RUN(launch_cqt_analysis_oct(n.tabs_adjS, n.fft, i, tp.spec, gTop[i], c.s));
    }
    std::vector<TV> gcat(no);
    for (int j = 0; j < no; ++j) gcat[j] = alloc_tv(c, 2 * cf.Ns[j], bins * (j + 1), Tof(j), true);
    // ---- decoder, last step first (unet.py:807-839) ----
    TV gXoutPrev;      // gradient with respect to the out-block accumulator input of the step processed before (i + 1)
    for (int i = no - 1; i >= 0; --i) {
        const int j = no - 1 - i, Fj = bins * (j + 1), Tj = Tof(j);
        const int dout = j == 0 ? cf.Ns[0] : cf.Ns[j - 1];
        TV gXout = alloc_tv(c, 2, Fj, Tj, false);
        RUN(launch_combine(gTop[i], TV(), 1.f, 0.f, slice_f(gXout, 0, bins), nullptr, c.s));
        if (j > 0) { RUN(launch_resample_up_adj(gXoutPrev, slice_f(gXout, bins, Fj - bins), 0.f, c.s)); c.release(gXoutPrev.p); }
        c.release(gTop[i].p);
        TV gXdec = alloc_tv(c, dout, Fj, Tj, true);
        resblock_bwd(c, n.ups_out[i], tp.ups_out[i], gXout, 0.5f, gXdec);
        RUN(launch_scale_channels(gXout, nullptr, 0, kInvSqrt2, gXout, c.s));     // -> gradient of the accumulator input
        gXoutPrev = gXout;
        if (j > 0) RUN(launch_resample_up_adj(slice_c(gcat[j - 1], 0, cf.Ns[j - 1]), slice_f(gXdec, bins, Fj - bins), 1.f, c.s));
        resblock_bwd(c, n.ups_main[i], tp.ups_main[i], gXdec, kInvSqrt2, gcat[j]);
        c.release(gXdec.p);
    }
    // ---- bottleneck (unet.py:800-804) ----
    const int Fl = bins * no, Tl = Tof(no - 1);
    TV gXm = slice_c(gcat[no - 1], 0, cf.Ns[no - 1]);
    resblock_bwd(c, n.mid_out, tp.mid_out, gXoutPrev, kInvSqrt2, gXm);
    c.release(gXoutPrev.p);
    TV gXmid = alloc_tv(c, cf.Ns[no - 1], Fl, Tl, true);
    resblock_bwd(c, n.mid_main, tp.mid_main, gXm, kInvSqrt2, gXmid);
    // ---- encoder, deepest level first (unet.py:747-795) ----
    TV gXcatNext, gpyrNext;
    for (int i = no - 1; i >= 0; --i) {
        const int Ti = Tof(i), Fi = bins * (i + 1);
        const int din = i == 0 ? cf.Ns[0] : cf.Ns[i - 1];
        TV gskip = slice_c(gcat[i], cf.Ns[i], cf.Ns[i]);
        const int Tpy = i < no - 1 ? Ti / 2 : Ti;
        TV gpyr = alloc_tv(c, 2, Fi, Tpy, true);
        if (i == no - 1) {    // Xmid = (pyr_down_proj(pyr) + skip) / sqrt 2
            RUN(launch_combine(gXmid, gskip, kInvSqrt2, 1.f, gskip, nullptr, c.s));
            conv_T(c, gXmid, n.downs[i].pyr, 1, gpyr, kInvSqrt2);
            c.release(gXmid.p);
        } else {              // Xcat_{i+1}[:, :, bins:] = (pyr_down_proj(pyr) + down(skip)) / sqrt 2
            TV gsl = slice_f(gXcatNext, bins, Fi);
            RUN(launch_scale_channels(gsl, nullptr, 0, kInvSqrt2, gsl, c.s));
            RUN(launch_resample_down_adj(gsl, gskip, 1.f, c.s));
            conv_T(c, gsl, n.downs[i].pyr, 1, gpyr, 1.f);
            // the pyramid of level i is also the lower rows of the next level's pyramid (down-sampled, or copied at the last level)
            TV src = slice_f(gpyrNext, bins, Fi);
            if (i + 1 < no - 1) RUN(launch_resample_down_adj(src, gpyr, 1.f, c.s));
            else RUN(launch_combine(src, gpyr, 1.f, 1.f, gpyr, nullptr, c.s));
            c.release(gXcatNext.p); c.release(gpyrNext.p);
        }
        TV gXcat = alloc_tv(c, din, Fi, Ti, true);
        resblock_bwd(c, n.downs[i].main, tp.main[i], gskip, kInvSqrt2, gXcat);
        c.release(gcat[i].p);
        TV gC = alloc_tv(c, 2, bins, Ti, true);
        resblock_bwd(c, n.downs[i].init, tp.init[i], slice_f(gXcat, 0, bins), kInvSqrt2, gC);
        if (i < no - 1) RUN(launch_resample_down_adj(slice_f(gpyr, 0, bins), gC, 1.f, c.s));
        else RUN(launch_combine(slice_f(gpyr, 0, bins), gC, 1.f, 1.f, gC, nullptr, c.s));
        RUN(launch_cqt_synth_oct(n.tabs_adjA, n.fft, no - 1 - i, gC, tp.Y, c.s));   // FFT_M of the coefficient gradients
        c.release(gC.p);
        gXcatNext = gXcat; gpyrNext = gpyr;
    }
    c.release(gXcatNext.p); c.release(gpyrNext.p);
    // ---- analysis adjoint (unet.py:743): spectrum gradient on the whole circle, then the adjoint of the real-input FFT ----
    RUN(launch_cqt_gather_adj(n.tabs_adjA, B, 0, no - 1, tp.Y, tp.spec, false, c.s));
    RUN(launch_fft_big(n.fft, B, true, nullptr, 0, 0.f, tp.spec, tp.tmp, tp.fscr, nullptr, g_x, L, tp.in_scale,
                       tp.skip_scale != 0.f ? g_out : nullptr, L, tp.skip_scale, c.s));
    if (!c.dry()) AID_CUDA_CHECK(cudaGetLastError());
}

static size_t plan_vjp(Net& n, int B) {
    Ctx c; c.n = &n; c.B = B; c.nsig = 1; c.ar.reset(nullptr, 0, true);
    Tape t; c.tape = &t;
    forward(c, nullptr, nullptr, nullptr, 1.f, 1.f, 0.f);
    backward(c, t, nullptr, nullptr);
    return c.ar.peak + 4096;
}

static size_t cqt_ws_bytes(const Net& n, int B) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return al((size_t)B * n.cfg.audio_len * sizeof(float2)) + 2 * al((size_t)B * n.fft.M * sizeof(float2)) +
           al((size_t)B * n.tabs.ytotal * sizeof(float2)) + 4096;
}

}  // namespace aid

// =======================================================================================================
using namespace aid;

struct aid_handle { Net net; };
static std::string g_create_error;

template <class Fn>
static int guarded(aid_handle* h, Fn&& fn) {
    try { fn(); return AID_OK; }
    catch (const CudaError& e) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e.code, cudaGetErrorString(e.code), e.file, e.line, e.expr);
        if (h) h->net.err = buf; else g_create_error = buf;
        return AID_ERR_CUDA;
    } catch (const std::invalid_argument& e) { if (h) h->net.err = e.what(); else g_create_error = e.what(); return AID_ERR_INVALID; }
    catch (const std::exception& e) {
        if (h) h->net.err = e.what(); else g_create_error = e.what();
        return std::string(e.what()).find("workspace") != std::string::npos ? AID_ERR_WORKSPACE : AID_ERR_STATE;
    }
}

extern "C" {

int aid_create(const aid_config* cfg, int device, aid_handle** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return AID_ERR_INVALID; }
    aid_handle* h = nullptr;
    int rc = guarded(nullptr, [&] {
        h = new aid_handle();
        h->net.cfg = *cfg; h->net.device = device;
        build_net(h->net);
        fill_host_tables(h->net);
        int slots = 0; plan_forward(h->net, 1, &slots);
        h->net.n_stat_slots = slots;
        // host-only so far: device tables are uploaded by aid_finalize / the first CQT call, so that the
        // schema (aid_weight_info) can be queried on a machine without a GPU
    });
    if (rc != AID_OK) { delete h; return rc; }
    *out = h;
    return AID_OK;
}

void aid_destroy(aid_handle* h) {
    if (!h) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->net.device);
    if (h->net.dweights) cudaFree(h->net.dweights);
    if (h->net.dweights_tc) cudaFree(h->net.dweights_tc);
    if (h->net.d_tables) cudaFree(h->net.d_tables);
    if (h->net.d_sat) cudaFree(h->net.d_sat);
    if (h->net.dweights_T) cudaFree(h->net.dweights_T);
    if (h->net.dweights_tcT) cudaFree(h->net.dweights_tcT);
    if (h->net.d_adj_tables) cudaFree(h->net.d_adj_tables);
    for (auto& r : h->net.prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto& e : h->net.prof_pool) cudaEventDestroy(e);
    if (prev >= 0) cudaSetDevice(prev);
    delete h;
}

const char* aid_last_error(const aid_handle* h) { return h ? h->net.err.c_str() : g_create_error.c_str(); }

int aid_num_weights(const aid_handle* h) { return h ? (int)h->net.weights.size() : 0; }

int aid_weight_info(const aid_handle* h, int index, const char** name, int64_t* shape4, int* ndim) {
    if (!h || index < 0 || index >= (int)h->net.weights.size()) return AID_ERR_INVALID;
    const Weight& w = h->net.weights[index];
    if (name) *name = w.name.c_str();
    if (ndim) *ndim = (int)w.shape.size();
    if (shape4) for (size_t i = 0; i < w.shape.size() && i < 4; ++i) shape4[i] = w.shape[i];
    return AID_OK;
}

int aid_load_weight(aid_handle* h, const char* name, const float* host, const int64_t* shape, int ndim) {
    if (!h || !name || !host || !shape) return AID_ERR_INVALID;
    return guarded(h, [&] {
        if (h->net.finalized) throw std::runtime_error("aid_load_weight after aid_finalize");
        auto it = h->net.windex.find(name);
        if (it == h->net.windex.end()) throw std::invalid_argument(std::string("unexpected key in state dict: ") + name);
        Weight& w = h->net.weights[it->second];
        if ((int)w.shape.size() != ndim) throw std::invalid_argument(std::string("rank mismatch for ") + name);
        for (int i = 0; i < ndim; ++i)
            if (w.shape[i] != shape[i]) throw std::invalid_argument(std::string("size mismatch for ") + name);
        if (!w.ignored) w.host.assign(host, host + w.numel());
        w.loaded = true;
    });
}

int aid_finalize(aid_handle* h) {
    if (!h) return AID_ERR_INVALID;
    return guarded(h, [&] { if (h->net.finalized) throw std::runtime_error("already finalized"); finalize_net(h->net); });
}

int aid_workspace_bytes(aid_handle* h, int B, size_t* bytes) {
    if (!h || !bytes || B < 1) return AID_ERR_INVALID;
    return guarded(h, [&] { *bytes = plan_forward(h->net, B, nullptr); });
}

int aid_unet_forward(aid_handle* h, const float* x_dev, const float* c_noise_dev, int n_sigma, float* out_dev, int B,
                     float in_scale, float out_scale, float skip_scale, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h || !x_dev || !c_noise_dev || !out_dev || !workspace_dev || B < 1) return AID_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->net.finalized) throw std::runtime_error("aid_unet_forward before aid_finalize");
        if (n_sigma != 1 && n_sigma != B) throw std::invalid_argument("n_sigma must be 1 or B");
        if (out_dev == x_dev && skip_scale != 0.f) throw std::invalid_argument("out may alias x only when skip_scale == 0");
        Ctx c; c.n = &h->net; c.B = B; c.nsig = n_sigma; c.s = (cudaStream_t)stream;
        c.ar.reset((char*)workspace_dev, workspace_bytes, false);
        forward(c, x_dev, c_noise_dev, out_dev, in_scale, out_scale, skip_scale);
    });
}

int aid_unet_forward_ds(aid_handle* h, const float* x_dev, const float* c_noise_dev, int n_sigma, float* out_dev, int B,
                        const float* scales_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h || !x_dev || !c_noise_dev || !out_dev || !workspace_dev || !scales_dev || B < 1) return AID_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->net.finalized) throw std::runtime_error("aid_unet_forward_ds before aid_finalize");
        if (n_sigma != 1 && n_sigma != B) throw std::invalid_argument("n_sigma must be 1 or B");
        if (out_dev == x_dev) throw std::invalid_argument("out may not alias x (the skip term reads x)");
        Ctx c; c.n = &h->net; c.B = B; c.nsig = n_sigma; c.s = (cudaStream_t)stream;
        c.ar.reset((char*)workspace_dev, workspace_bytes, false);
        forward(c, x_dev, c_noise_dev, out_dev, 1.f, 1.f, 0.f, scales_dev);
    });
}

int aid_vjp_workspace_bytes(aid_handle* h, int B, size_t* bytes) {
    if (!h || !bytes || B < 1) return AID_ERR_INVALID;
    return guarded(h, [&] {
        if (!h->net.finalized) throw std::runtime_error("aid_vjp_workspace_bytes before aid_finalize");
        *bytes = plan_vjp(h->net, B);
    });
}

int aid_unet_forward_tape(aid_handle* h, const float* x_dev, const float* c_noise_dev, int n_sigma, float* out_dev, int B,
                          float in_scale, float out_scale, float skip_scale, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!h || !x_dev || !c_noise_dev || !out_dev || !workspace_dev || B < 1) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net;
        if (!n.finalized) throw std::runtime_error("aid_unet_forward_tape before aid_finalize");
        if (n_sigma != 1 && n_sigma != B) throw std::invalid_argument("n_sigma must be 1 or B");
        if (out_dev == x_dev) throw std::invalid_argument("out may not alias x on the taped path");
        ensure_vjp(n);
        n.tape.valid = false;
        Ctx c; c.n = &n; c.B = B; c.nsig = n_sigma; c.s = (cudaStream_t)stream;
        c.ar.reset((char*)workspace_dev, workspace_bytes, false);
        c.tape = &n.tape;
        forward(c, x_dev, c_noise_dev, out_dev, in_scale, out_scale, skip_scale);
        n.tape.B = B; n.tape.nsig = n_sigma; n.tape.in_scale = in_scale; n.tape.out_scale = out_scale; n.tape.skip_scale = skip_scale;
        n.tape.ws = (char*)workspace_dev; n.tape.ws_bytes = workspace_bytes;
        n.tape_arena = c.ar;
        ++n.tape.generation;
        n.tape.valid = true;
    });
}

int aid_unet_backward(aid_handle* h, const float* grad_out_dev, float* grad_x_dev, void* stream) {
    if (!h || !grad_out_dev || !grad_x_dev) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net;
        if (!n.tape.valid) throw std::runtime_error("aid_unet_backward without a live tape (call aid_unet_forward_tape first; its workspace must stay untouched)");
        if (grad_out_dev == grad_x_dev) throw std::invalid_argument("grad_x may not alias grad_out");
        Ctx c; c.n = &n; c.B = n.tape.B; c.nsig = n.tape.nsig; c.s = (cudaStream_t)stream;
        c.ar = n.tape_arena;       // a copy: the tape's own allocations stay, everything the backward takes is returned
        c.mod = n.tape.mod;
        backward(c, n.tape, grad_out_dev, grad_x_dev);
    });
}

int aid_cqt_layout(const aid_handle* h, int B, int64_t* offsets, int32_t* frames) {
    if (!h || !offsets) return AID_ERR_INVALID;
    const Net& n = h->net;
    int64_t off = 0;
    for (int o = 0; o < n.cfg.num_octs; ++o) {
        offsets[o] = off;
        if (frames) frames[o] = n.tabs.M[o];
        off += (int64_t)B * 2 * n.cfg.bins_per_oct * n.tabs.M[o];
    }
    offsets[n.cfg.num_octs] = off;
    return AID_OK;
}

int aid_cqt_plan(const aid_handle* h, int32_t* K, int32_t* n_win, int32_t* centre, int32_t* Lg, int32_t* woff, float* win,
                 float* dual, float* hhpf) {
    if (!h) return AID_ERR_INVALID;
    const CqtPlanHost& p = h->net.plan;
    if (K) *K = p.K;
    if (n_win) *n_win = (int32_t)p.win.size();
    if (centre) std::copy(p.centre.begin(), p.centre.end(), centre);
    if (Lg) std::copy(p.Lg.begin(), p.Lg.end(), Lg);
    if (woff) std::copy(p.woff.begin(), p.woff.end(), woff);
    if (win) std::copy(p.win.begin(), p.win.end(), win);
    if (dual) std::copy(p.dual.begin(), p.dual.end(), dual);
    if (hhpf) std::copy(p.hhpf.begin(), p.hhpf.end(), hhpf);
    return AID_OK;
}

int aid_cqt_workspace_bytes(const aid_handle* h, int B, size_t* bytes) {
    if (!h || !bytes || B < 1) return AID_ERR_INVALID;
    *bytes = cqt_ws_bytes(h->net, B);
    return AID_OK;
}

static void cqt_ws_split(Net& n, int B, void* ws, size_t ws_bytes, float2** spec, float2** tmp, float2** fscr, float2** Y) {
    upload_tables(n);
    if (ws_bytes < cqt_ws_bytes(n, B)) throw std::runtime_error("workspace too small");
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char* p = (char*)ws;
    *spec = (float2*)p; p += al((size_t)B * n.cfg.audio_len * sizeof(float2));
    *tmp = (float2*)p; p += al((size_t)B * n.fft.M * sizeof(float2));
    *fscr = (float2*)p; p += al((size_t)B * n.fft.M * sizeof(float2));
    *Y = (float2*)p;
}

int aid_cqt_fwd(aid_handle* h, const float* x_dev, float* coef_dev, int B, void* ws, size_t ws_bytes, void* stream) {
    if (!h || !x_dev || !coef_dev || !ws) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net; cudaStream_t s = (cudaStream_t)stream;
        float2 *spec, *tmp, *fscr, *Y; cqt_ws_split(n, B, ws, ws_bytes, &spec, &tmp, &fscr, &Y);
        const int L = n.cfg.audio_len, bins = n.cfg.bins_per_oct;
        launch_fft_big(n.fft, B, false, x_dev, L, 1.f, nullptr, tmp, fscr, spec, nullptr, 0, 0.f, nullptr, 0, 0.f, s);
        int64_t off = 0;
        for (int o = 0; o < n.cfg.num_octs; ++o) {
            TV C = make_tv(coef_dev + off, B, 2, bins, n.tabs.M[o]);
            launch_cqt_analysis_oct(n.tabs, n.fft, o, spec, C, s);
            off += (int64_t)B * 2 * bins * n.tabs.M[o];
        }
        AID_CUDA_CHECK(cudaGetLastError());
    });
}

int aid_cqt_bwd(aid_handle* h, const float* coef_dev, float* x_dev, int B, void* ws, size_t ws_bytes, void* stream) {
    if (!h || !x_dev || !coef_dev || !ws) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net; cudaStream_t s = (cudaStream_t)stream;
        float2 *spec, *tmp, *fscr, *Y; cqt_ws_split(n, B, ws, ws_bytes, &spec, &tmp, &fscr, &Y);
        const int L = n.cfg.audio_len, bins = n.cfg.bins_per_oct;
        int64_t off = 0;
        for (int o = 0; o < n.cfg.num_octs; ++o) {
            TV C = make_tv(const_cast<float*>(coef_dev) + off, B, 2, bins, n.tabs.M[o]);
            launch_cqt_synth_oct(n.tabs, n.fft, o, C, Y, s);
            off += (int64_t)B * 2 * bins * n.tabs.M[o];
        }
        launch_cqt_synth_gather(n.tabs, B, Y, spec, s);
        launch_fft_big(n.fft, B, true, nullptr, 0, 0.f, spec, tmp, fscr, nullptr, x_dev, L, 1.f / (float)L, nullptr, 0, 0.f, s);
        AID_CUDA_CHECK(cudaGetLastError());
    });
}

int aid_hpf_dc(aid_handle* h, const float* x_dev, float* out_dev, int B, void* ws, size_t ws_bytes, void* stream) {
    if (!h || !x_dev || !out_dev || !ws) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net; cudaStream_t s = (cudaStream_t)stream;
        float2 *spec, *tmp, *fscr, *Y; cqt_ws_split(n, B, ws, ws_bytes, &spec, &tmp, &fscr, &Y);
        const int L = n.cfg.audio_len;
        launch_fft_big(n.fft, B, false, x_dev, L, 1.f, nullptr, tmp, fscr, spec, nullptr, 0, 0.f, nullptr, 0, 0.f, s);
        launch_spec_mul_real(B, L, spec, n.tabs.hhpf, s);
        launch_fft_big(n.fft, B, true, nullptr, 0, 0.f, spec, tmp, fscr, nullptr, out_dev, L, 1.f / (float)L, nullptr, 0, 0.f, s);
        AID_CUDA_CHECK(cudaGetLastError());
    });
}

int aid_edm_add_noise(float* x_dev, const float* eps_dev, float scale, int64_t n, void* stream) {
    if (!x_dev || !eps_dev || n < 0) return AID_ERR_INVALID;
    launch_axpy_noise(x_dev, eps_dev, scale, n, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_edm_step(const float* xin, const float* xhat, const float* y, const float* mask, int64_t mask_n, int64_t n, float sigma,
                 float hstep, int mode, const float* d_prev, const float* xbase, float* d_out, float* x_out, void* stream) {
    if (!xin || !xhat || !x_out || n < 0 || (mask && (!y || mask_n <= 0)) || (mode == 1 && (!d_prev || !xbase)) || (mode != 0 && mode != 1))
        return AID_ERR_INVALID;
    launch_edm_step(xin, xhat, y, mask, mask_n, n, sigma, hstep, mode, d_prev, xbase, d_out, x_out, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_edm_step_ds(const float* xin, const float* xhat, const float* y, const float* mask, int64_t mask_n, int64_t n,
                    const float* sigma_h_dev, int mode, const float* d_prev, const float* xbase, float* d_out, float* x_out, void* stream) {
    if (!xin || !xhat || !x_out || !sigma_h_dev || n < 0 || (mask && (!y || mask_n <= 0)) || (mode == 1 && (!d_prev || !xbase)) ||
        (mode != 0 && mode != 1))
        return AID_ERR_INVALID;
    launch_edm_step(xin, xhat, y, mask, mask_n, n, 1.f, 0.f, mode, d_prev, xbase, d_out, x_out, (cudaStream_t)stream, sigma_h_dev);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_philox_normal(float* x_dev, int n_clips, int64_t L, uint64_t seed, uint32_t stream_id, uint32_t clip0, uint32_t draw, float scale,
                      int accumulate, const float* scale_draw_dev, void* stream) {
    if (!x_dev || n_clips < 0 || L < 0 || L > (1ll << 33)) return AID_ERR_INVALID;
    launch_philox_normal(x_dev, n_clips, L, seed, stream_id, clip0, draw, scale, accumulate != 0, scale_draw_dev, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_sched_select(const float* table_dev, int row_floats, int32_t* counter_dev, float* cur_dev, void* stream) {
    if (!table_dev || !counter_dev || !cur_dev || row_floats < 1 || row_floats > 64) return AID_ERR_INVALID;
    launch_sched_select(table_dev, row_floats, counter_dev, cur_dev, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_spectral_mask(const float* x, const float* y, const float* mask, int B, int64_t L, int n_fft, int hop, int n_frames, float* frames,
                      size_t frames_bytes, float* out, void* stream) {
    if (!x || !mask || !frames || !out || B <= 0 || L <= 0 || L > (1ll << 30) || n_fft < 4 || n_fft > 4096 || (n_fft & (n_fft - 1)) != 0 ||
        hop <= 0 || hop > n_fft)
        return AID_ERR_INVALID;
    const int64_t Lp = L + (n_fft - L % n_fft);
    if (n_frames != 1 + Lp / hop || frames_bytes < (size_t)B * n_frames * n_fft * sizeof(float)) return AID_ERR_INVALID;
    try {
        launch_spectral_mask(x, y, mask, B, (int)L, n_fft, hop, n_frames, frames, out, (cudaStream_t)stream);
    } catch (const CudaError&) {
        return AID_ERR_CUDA;
    }
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

// ---- single-operator entry points -----------------------------------------------------------------------
static int op_conv2d_impl(const float* a_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                          const float* gate_dev, const float* R_dev, const float* R2_dev, float alpha, float beta, float* out_dev,
                          double* stats_dev, int mode, void* stream, int iters, float* ms_out) {
    if (!a_dev || !w_dev || !out_dev) return AID_ERR_INVALID;
    try {
        cudaStream_t s = (cudaStream_t)stream;
        float* wp = nullptr;
        const size_t e = (size_t)Cout * Cin * KF * KT;
        AID_CUDA_CHECK(cudaMalloc(&wp, e * sizeof(float)));
        pack_conv_weight_kernel<<<(int)std::min<size_t>(4096, (e + 255) / 256), 256, 0, s>>>(w_dev, wp, Cout, Cin, KF * KT);
        TV a = make_tv(const_cast<float*>(a_dev), B, Cin, F, T), out = make_tv(out_dev, B, Cout, F, T);
        ConvEpilogue ep; ep.gate = gate_dev; ep.gate_bstride = 0; ep.alpha = alpha; ep.beta = beta; ep.stats = stats_dev;
        if (R_dev) ep.R = make_tv(const_cast<float*>(R_dev), B, Cout, F, T);
        if (R2_dev) ep.R2 = make_tv(const_cast<float*>(R2_dev), B, Cout, F, T);
        cudaEvent_t e0, e1;
        AID_CUDA_CHECK(cudaEventCreate(&e0)); AID_CUDA_CHECK(cudaEventCreate(&e1));
        __half *wtc = nullptr, *ah = nullptr;
        const int pf = tc_pad_rows(T, KF, dil);
        const size_t ahalves = (size_t)B * Cin * (F + 2 * pf) * (T + 2);
        int sms = 148, dev = 0;
        if (mode == 1 || mode == 3 || mode == 4) {
            if (!conv_tc_supported(Cin, Cout, KF, KT) || R2_dev) throw std::invalid_argument("shape not supported by the tcgen05 path");
            AID_CUDA_CHECK(cudaGetDevice(&dev));
            AID_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            if (mode == 3 || mode == 4) {
                AID_CUDA_CHECK(cudaMalloc(&wtc, tc2_weight_halves(Cout, Cin, KF, KT) * sizeof(__half)));
                AID_CUDA_CHECK(cudaMalloc(&ah, tc2_act_halves(B, Cin, F, T, pf) * sizeof(__half)));
                launch_pack_weight_tc2(w_dev, wtc, Cout, Cin, KF, KT, s);
                if (mode == 4) {   // a, R and out are channels-last [B][F][T][C]
                    launch_gn_act_tc2_cl(a_dev, B, Cin, F, T, nullptr, 1, nullptr, nullptr, 0, false, pf, ah, s);
                    out = make_tv_cl(out_dev, B, Cout, F, T);
                    if (R_dev) ep.R = make_tv_cl(const_cast<float*>(R_dev), B, Cout, F, T);
                    ep.R_cl = R_dev != nullptr; ep.out_cl = true;
                } else launch_to_planar_tc2(a, pf, ah, s);
            } else {
                AID_CUDA_CHECK(cudaMalloc(&wtc, 2 * e * sizeof(__half)));
                AID_CUDA_CHECK(cudaMalloc(&ah, 2 * ahalves * sizeof(__half)));
                launch_pack_weight_tc(w_dev, wtc, Cout, Cin, KF, KT, 2, s);
                launch_to_planar_tc(a, pf, ah, ah + ahalves, s);
            }
        } else if (mode != 0 && mode != 2) throw std::invalid_argument("unknown conv mode");
        (void)iters;
        AID_CUDA_CHECK(cudaEventRecord(e0, s));
        if (mode == 3 || mode == 4) launch_conv_tc2(ah, pf, wtc, B, Cin, F, T, KF, KT, dil, out, ep, sms, s);
        else if (mode == 1) launch_conv_tc(ah, ah + ahalves, pf, wtc, B, Cin, F, T, KF, KT, dil, out, ep, sms, s);
        else if (mode == 2 || !launch_conv_thin(a, wp, KF, KT, dil, out, ep, s)) launch_conv_simt(a, wp, KF, KT, dil, out, ep, s);
        AID_CUDA_CHECK(cudaEventRecord(e1, s));
        AID_CUDA_CHECK(cudaGetLastError());
        AID_CUDA_CHECK(cudaStreamSynchronize(s));
        if (ms_out) AID_CUDA_CHECK(cudaEventElapsedTime(ms_out, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (wtc) AID_CUDA_CHECK(cudaFree(wtc));
        if (ah) AID_CUDA_CHECK(cudaFree(ah));
        AID_CUDA_CHECK(cudaFree(wp));
        return AID_OK;
    } catch (const CudaError& e) { fprintf(stderr, "aid_op_conv2d: CUDA error %s at %s:%d\n", cudaGetErrorString(e.code), e.file, e.line); return AID_ERR_CUDA; }
    catch (const std::exception& e) { fprintf(stderr, "aid_op_conv2d: %s\n", e.what()); return AID_ERR_INVALID; }
}

int aid_op_conv2d(const float* a_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                  const float* gate_dev, const float* R_dev, const float* R2_dev, float alpha, float beta, float* out_dev,
                  double* stats_dev, int mode, void* stream) {
    return op_conv2d_impl(a_dev, w_dev, B, Cin, Cout, F, T, KF, KT, dil, gate_dev, R_dev, R2_dev, alpha, beta, out_dev, stats_dev, mode,
                          stream, 1, nullptr);
}

/* debug/tuning: like aid_op_conv2d but returns the device time of the convolution kernel alone (one launch, after a
 * warm-up launch); not part of the drop-in surface */
int aid_debug_time_conv2d(const float* a_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                          const float* gate_dev, const float* R_dev, float alpha, float* out_dev, double* stats_dev, int mode,
                          float* ms_out) {
    int rc = op_conv2d_impl(a_dev, w_dev, B, Cin, Cout, F, T, KF, KT, dil, gate_dev, R_dev, nullptr, alpha, 0.f, out_dev, stats_dev, mode,
                            nullptr, 1, nullptr);
    if (rc != AID_OK) return rc;
    return op_conv2d_impl(a_dev, w_dev, B, Cin, Cout, F, T, KF, KT, dil, gate_dev, R_dev, nullptr, alpha, 0.f, out_dev, stats_dev, mode,
                          nullptr, 1, ms_out);
}

/* debug / layout parity: the conv_mode 2 operand layouts.  a_out: tc2_act_halves fp16, w_out: tc2_weight_halves fp16 (either may be NULL) */
int aid_debug_tc2_operands(const float* x_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int PF,
                           void* a_out_dev, void* w_out_dev, uint64_t* a_halves, uint64_t* w_halves) {
    if (a_halves) *a_halves = tc2_act_halves(B, Cin, F, T, PF);
    if (w_halves) *w_halves = tc2_weight_halves(Cout, Cin, KF, KT);
    try {
        if (a_out_dev && x_dev) launch_to_planar_tc2(make_tv(const_cast<float*>(x_dev), B, Cin, F, T), PF, (__half*)a_out_dev, nullptr);
        if (w_out_dev && w_dev) launch_pack_weight_tc2(w_dev, (__half*)w_out_dev, Cout, Cin, KF, KT, nullptr);
        AID_CUDA_CHECK(cudaGetLastError());
        AID_CUDA_CHECK(cudaDeviceSynchronize());
        return AID_OK;
    } catch (const CudaError& e) { fprintf(stderr, "aid_debug_tc2_operands: CUDA error %s\n", cudaGetErrorString(e.code)); return AID_ERR_CUDA; }
}

/* debug / tuning: device time (ms per launch, averaged over iters) of the conv_mode 2 normalise + GELU + operand-layout pass */
int aid_debug_time_gn_tc2(const float* x_dev, int B, int C, int F, int T, int PF, int iters, float* ms_out) {
    try {
        TV x = make_tv(const_cast<float*>(x_dev), B, C, F, T);
        double* st = nullptr; float *gamma = nullptr, *aff = nullptr; __half* a = nullptr;
        AID_CUDA_CHECK(cudaMalloc(&st, (size_t)B * 16 * sizeof(double)));
        AID_CUDA_CHECK(cudaMalloc(&gamma, C * sizeof(float))); AID_CUDA_CHECK(cudaMalloc(&aff, C * sizeof(float)));
        AID_CUDA_CHECK(cudaMemset(gamma, 0, C * sizeof(float))); AID_CUDA_CHECK(cudaMemset(aff, 0, C * sizeof(float)));
        AID_CUDA_CHECK(cudaMalloc(&a, tc2_act_halves(B, C, F, T, PF) * sizeof(__half)));
        AID_CUDA_CHECK(cudaMemset(st, 0, (size_t)B * 16 * sizeof(double)));
        launch_group_stats(x, st, nullptr);
        cudaEvent_t e0, e1; AID_CUDA_CHECK(cudaEventCreate(&e0)); AID_CUDA_CHECK(cudaEventCreate(&e1));
        launch_gn_act_tc2(x, st, (long long)(C / 8) * F * T, gamma, aff, 0, true, PF, a, nullptr);
        AID_CUDA_CHECK(cudaEventRecord(e0, nullptr));
        for (int i = 0; i < iters; ++i) launch_gn_act_tc2(x, st, (long long)(C / 8) * F * T, gamma, aff, 0, true, PF, a, nullptr);
        AID_CUDA_CHECK(cudaEventRecord(e1, nullptr));
        AID_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f; AID_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_out) *ms_out = ms / iters;
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(st); cudaFree(gamma); cudaFree(aff); cudaFree(a);
        return AID_OK;
    } catch (const CudaError& e) { fprintf(stderr, "aid_debug_time_gn_tc2: CUDA error %s\n", cudaGetErrorString(e.code)); return AID_ERR_CUDA; }
}

/* debug / parity / tuning: one dilated residual layer of conv_mode 2, two-kernel path (fused = 0) or conv_comb_kernel (fused = 1) */
int aid_debug_dilated_layer(const float* x_dev, const float* w_dev, int B, int C, int F, int T, int dil, const float* gamma_dev,
                            const float* affine_dev, const float* gate_dev, float alpha, int fused, float* out_dev, double* stats_out_dev,
                            float* ms_out) {
    if (!x_dev || !w_dev || !gamma_dev || !out_dev || C % 8 != 0) return AID_ERR_INVALID;
    try {
        if (!conv_tc_supported(C, C, 5, 3) || (fused && !conv_comb_supported(C, F, T, dil))) throw std::invalid_argument("shape not supported");
        TV x = make_tv(const_cast<float*>(x_dev), B, C, F, T), out = make_tv(out_dev, B, C, F, T);
        int sms = 148, dev = 0;
        AID_CUDA_CHECK(cudaGetDevice(&dev));
        AID_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int pf = tc_pad_rows(T, 5, dil);
        double* st = nullptr; __half *wtc = nullptr, *a = nullptr;
        AID_CUDA_CHECK(cudaMalloc(&st, (size_t)B * 16 * sizeof(double)));
        AID_CUDA_CHECK(cudaMemset(st, 0, (size_t)B * 16 * sizeof(double)));
        AID_CUDA_CHECK(cudaMalloc(&wtc, tc2_weight_halves(C, C, 5, 3) * sizeof(__half)));
        AID_CUDA_CHECK(cudaMalloc(&a, tc2_act_halves(B, C, F, T, pf) * sizeof(__half)));
        launch_pack_weight_tc2(w_dev, wtc, C, C, 5, 3, nullptr);
        __half* wcomb = nullptr;
        if (fused && C == 96) {
            AID_CUDA_CHECK(cudaMalloc(&wcomb, comb_weight_halves(96) * sizeof(__half)));
            launch_pack_weight_comb(w_dev, wcomb, 96, nullptr);
        }
        launch_group_stats(x, st, nullptr);
        const long long n_grp = (long long)(C / 8) * F * T;
        ConvEpilogue ep; ep.gate = gate_dev; ep.gate_bstride = 0; ep.alpha = alpha; ep.R = x; ep.stats = stats_out_dev;
        cudaEvent_t e0, e1; AID_CUDA_CHECK(cudaEventCreate(&e0)); AID_CUDA_CHECK(cudaEventCreate(&e1));
        for (int rep = 0; rep < (ms_out ? 2 : 1); ++rep) {
            if (stats_out_dev) AID_CUDA_CHECK(cudaMemsetAsync(stats_out_dev, 0, (size_t)B * 16 * sizeof(double), nullptr));
            AID_CUDA_CHECK(cudaEventRecord(e0, nullptr));
            if (fused) launch_conv_comb(x, st, n_grp, gamma_dev, affine_dev, 0, wcomb ? wcomb : wtc, dil, out, ep, sms, nullptr);
            else {
                launch_gn_act_tc2(x, st, n_grp, gamma_dev, affine_dev, 0, true, pf, a, nullptr);
                launch_conv_tc2(a, pf, wtc, B, C, F, T, 5, 3, dil, out, ep, sms, nullptr);
            }
            AID_CUDA_CHECK(cudaEventRecord(e1, nullptr));
            AID_CUDA_CHECK(cudaGetLastError());
            AID_CUDA_CHECK(cudaEventSynchronize(e1));
        }
        if (ms_out) AID_CUDA_CHECK(cudaEventElapsedTime(ms_out, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(st); cudaFree(wtc); cudaFree(a); if (wcomb) cudaFree(wcomb);
        return AID_OK;
    } catch (const CudaError& e) { fprintf(stderr, "aid_debug_dilated_layer: CUDA error %s at %s:%d\n", cudaGetErrorString(e.code), e.file, e.line); return AID_ERR_CUDA; }
    catch (const std::exception& e) { fprintf(stderr, "aid_debug_dilated_layer: %s\n", e.what()); return AID_ERR_INVALID; }
}

/* debug / parity / tuning: the init block of an encoder level in conv_mode 2 (unet.py:452-493 with dim = 2, one 1x1 layer):
   fused = 1: conv_init.cu; fused = 0: the five un-fused launches of resblock().  Weights in the checkpoint layout
   (w_in, w_res: [N][2], wH: [N][N]); gamma, affine, gate: [N] (shared by the batch). */
int aid_debug_init_block(const float* x2_dev, const float* w_in_dev, const float* w_res_dev, const float* wH_dev, int B, int N, int F, int T,
                         const float* gamma_dev, const float* affine_dev, const float* gate_dev, int fused, float* out_dev,
                         double* stats_out_dev, float* ms_out) {
    if (!x2_dev || !w_in_dev || !w_res_dev || !wH_dev || !gamma_dev || !out_dev || N % 8 != 0) return AID_ERR_INVALID;
    try {
        if (!conv_tc_supported(N, N, 1, 1) || (fused && !init_block_supported(N, T))) throw std::invalid_argument("shape not supported");
        TV x2 = make_tv(const_cast<float*>(x2_dev), B, 2, F, T), out = make_tv(out_dev, B, N, F, T);
        int sms = 148, dev = 0;
        AID_CUDA_CHECK(cudaGetDevice(&dev));
        AID_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int pf = tc_pad_rows(T, 1, 1);
        const size_t plane = (size_t)B * N * F * T;
        double* st = nullptr; __half *wtc = nullptr, *a = nullptr; float *wk = nullptr, *scratch = nullptr, *y = nullptr, *x = nullptr;
        AID_CUDA_CHECK(cudaMalloc(&st, (size_t)B * 16 * sizeof(double)));
        AID_CUDA_CHECK(cudaMalloc(&wtc, tc2_weight_halves(N, N, 1, 1) * sizeof(__half)));
        AID_CUDA_CHECK(cudaMalloc(&wk, (size_t)4 * N * sizeof(float)));
        AID_CUDA_CHECK(cudaMalloc(&scratch, init_block_scratch_floats(B, N) * sizeof(float)));
        if (!fused) {
            AID_CUDA_CHECK(cudaMalloc(&a, tc2_act_halves(B, N, F, T, pf) * sizeof(__half)));
            AID_CUDA_CHECK(cudaMalloc(&y, plane * sizeof(float)));
            AID_CUDA_CHECK(cudaMalloc(&x, plane * sizeof(float)));
        }
        launch_pack_weight_tc2(wH_dev, wtc, N, N, 1, 1, nullptr);
        pack_conv_weight_kernel<<<1, 256>>>(w_in_dev, wk, N, 2, 1);
        pack_conv_weight_kernel<<<1, 256>>>(w_res_dev, wk + 2 * N, N, 2, 1);
        const long long n_grp = (long long)(N / 8) * F * T;
        cudaEvent_t e0, e1; AID_CUDA_CHECK(cudaEventCreate(&e0)); AID_CUDA_CHECK(cudaEventCreate(&e1));
        for (int rep = 0; rep < (ms_out ? 2 : 1); ++rep) {
            if (stats_out_dev) AID_CUDA_CHECK(cudaMemsetAsync(stats_out_dev, 0, (size_t)B * 16 * sizeof(double), nullptr));
            AID_CUDA_CHECK(cudaMemsetAsync(st, 0, (size_t)B * 16 * sizeof(double), nullptr));
            AID_CUDA_CHECK(cudaEventRecord(e0, nullptr));
            if (fused) {
                launch_init_block(x2, wk, wk + 2 * N, wtc, gamma_dev, affine_dev, 0, gate_dev, 0, out, stats_out_dev, scratch, sms, nullptr);
            } else {
                TV yv = make_tv(y, B, N, F, T), xv = make_tv(x, B, N, F, T);
                ConvEpilogue e1p; e1p.stats = st;
                if (!launch_conv_thin(x2, wk, 1, 1, 1, yv, e1p, nullptr)) throw std::invalid_argument("thin conv not applicable");
                launch_gn_act_tc2(yv, st, n_grp, gamma_dev, affine_dev, 0, true, pf, a, nullptr);
                ConvEpilogue e2p; e2p.gate = gate_dev; e2p.gate_bstride = 0; e2p.alpha = kInvSqrt2; e2p.R = yv;
                launch_conv_tc2(a, pf, wtc, B, N, F, T, 1, 1, 1, xv, e2p, sms, nullptr);
                ConvEpilogue e3p; e3p.R = xv; e3p.alpha = kInvSqrt2; e3p.stats = stats_out_dev;
                if (!launch_conv_thin(x2, wk + 2 * N, 1, 1, 1, out, e3p, nullptr)) throw std::invalid_argument("thin conv not applicable");
            }
            AID_CUDA_CHECK(cudaEventRecord(e1, nullptr));
            AID_CUDA_CHECK(cudaGetLastError());
            AID_CUDA_CHECK(cudaEventSynchronize(e1));
        }
        if (ms_out) AID_CUDA_CHECK(cudaEventElapsedTime(ms_out, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(st); cudaFree(wtc); cudaFree(wk); cudaFree(scratch);
        if (a) cudaFree(a); if (y) cudaFree(y); if (x) cudaFree(x);
        return AID_OK;
    } catch (const CudaError& e) { fprintf(stderr, "aid_debug_init_block: CUDA error %s at %s:%d\n", cudaGetErrorString(e.code), e.file, e.line); return AID_ERR_CUDA; }
    catch (const std::exception& e) { fprintf(stderr, "aid_debug_init_block: %s\n", e.what()); return AID_ERR_INVALID; }
}

/* debug / parity / tuning: an out block in conv_mode 2 (unet.py:452-493 with one 1x1 layer, proj_out and res_conv N -> 2; accum_dev != NULL:
   out = (accum + block(x)) / sqrt 2, unet.py:817).  fused = 1: out_block.cu; fused = 0: the four un-fused launches of resblock().
   Weights in the checkpoint layout (wH: [N][N], wP, wR: [2][N]); gamma, affine, gate: [N]. */
int aid_debug_out_block(const float* x_dev, const float* wH_dev, const float* wP_dev, const float* wR_dev, int B, int N, int F, int T,
                        const float* gamma_dev, const float* affine_dev, const float* gate_dev, const float* accum_dev, int fused, float* out_dev,
                        float* ms_out) {
    if (!x_dev || !wH_dev || !wP_dev || !wR_dev || !gamma_dev || !out_dev || N % 8 != 0) return AID_ERR_INVALID;
    try {
        TV x = make_tv(const_cast<float*>(x_dev), B, N, F, T), out = make_tv(out_dev, B, 2, F, T);
        TV acc = accum_dev ? make_tv(const_cast<float*>(accum_dev), B, 2, F, T) : TV();
        if (!conv_tc_supported(N, N, 1, 1) || (fused && !out_block_supported(x, out, acc))) throw std::invalid_argument("shape not supported");
        int sms = 148, dev = 0;
        AID_CUDA_CHECK(cudaGetDevice(&dev));
        AID_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int pf = tc_pad_rows(T, 1, 1);
        const size_t plane = (size_t)B * N * F * T;
        double* st = nullptr; __half *wtc = nullptr, *a = nullptr; float *wk = nullptr, *scratch = nullptr, *x1 = nullptr, *t = nullptr;
        AID_CUDA_CHECK(cudaMalloc(&st, (size_t)B * 16 * sizeof(double)));
        AID_CUDA_CHECK(cudaMemset(st, 0, (size_t)B * 16 * sizeof(double)));
        AID_CUDA_CHECK(cudaMalloc(&wtc, tc2_weight_halves(N, N, 1, 1) * sizeof(__half)));
        AID_CUDA_CHECK(cudaMalloc(&wk, ((size_t)N * N + 4 * N) * sizeof(float)));
        AID_CUDA_CHECK(cudaMalloc(&scratch, out_block_scratch_floats(B, N) * sizeof(float)));
        if (!fused) {
            AID_CUDA_CHECK(cudaMalloc(&a, tc2_act_halves(B, N, F, T, pf) * sizeof(__half)));
            AID_CUDA_CHECK(cudaMalloc(&x1, plane * sizeof(float)));
            AID_CUDA_CHECK(cudaMalloc(&t, (size_t)B * 2 * F * T * sizeof(float)));
        }
        float *hk = wk, *pk = wk + (size_t)N * N, *rk = pk + 2 * N;
        launch_pack_weight_tc2(wH_dev, wtc, N, N, 1, 1, nullptr);
        pack_conv_weight_kernel<<<64, 256>>>(wH_dev, hk, N, N, 1);
        pack_conv_weight_kernel<<<1, 256>>>(wP_dev, pk, 2, N, 1);
        pack_conv_weight_kernel<<<1, 256>>>(wR_dev, rk, 2, N, 1);
        launch_group_stats(x, st, nullptr);
        const long long n_grp = (long long)(N / 8) * F * T;
        cudaEvent_t e0, e1; AID_CUDA_CHECK(cudaEventCreate(&e0)); AID_CUDA_CHECK(cudaEventCreate(&e1));
        for (int rep = 0; rep < (ms_out ? 2 : 1); ++rep) {
            if (acc.p && rep > 0) break;     // accum is read per run; with out aliasing it a second run would accumulate twice (timing: pass accum = NULL)
            AID_CUDA_CHECK(cudaEventRecord(e0, nullptr));
            if (fused) {
                launch_out_block(x, st, n_grp, gamma_dev, affine_dev, 0, gate_dev, 0, hk, pk, rk, out, acc, scratch, nullptr);
            } else {
                TV x1v = make_tv(x1, B, N, F, T), tv = make_tv(t, B, 2, F, T);
                launch_gn_act_tc2(x, st, n_grp, gamma_dev, affine_dev, 0, true, pf, a, nullptr);
                ConvEpilogue e1p; e1p.gate = gate_dev; e1p.gate_bstride = 0; e1p.alpha = kInvSqrt2; e1p.R = x;
                launch_conv_tc2(a, pf, wtc, B, N, F, T, 1, 1, 1, x1v, e1p, sms, nullptr);
                if (!launch_conv_thin(x1v, pk, 1, 1, 1, tv, ConvEpilogue(), nullptr)) throw std::invalid_argument("thin conv not applicable");
                ConvEpilogue e3p; e3p.R = tv;
                if (acc.p) { e3p.alpha = 0.5f; e3p.beta = kInvSqrt2; e3p.R2 = acc; } else e3p.alpha = kInvSqrt2;
                if (!launch_conv_thin(x, rk, 1, 1, 1, out, e3p, nullptr)) throw std::invalid_argument("thin conv not applicable");
            }
            AID_CUDA_CHECK(cudaEventRecord(e1, nullptr));
            AID_CUDA_CHECK(cudaGetLastError());
            AID_CUDA_CHECK(cudaEventSynchronize(e1));
        }
        if (ms_out) AID_CUDA_CHECK(cudaEventElapsedTime(ms_out, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(st); cudaFree(wtc); cudaFree(wk); cudaFree(scratch);
        if (a) cudaFree(a); if (x1) cudaFree(x1); if (t) cudaFree(t);
        return AID_OK;
    } catch (const CudaError& e) { fprintf(stderr, "aid_debug_out_block: CUDA error %s at %s:%d\n", cudaGetErrorString(e.code), e.file, e.line); return AID_ERR_CUDA; }
    catch (const std::exception& e) { fprintf(stderr, "aid_debug_out_block: %s\n", e.what()); return AID_ERR_INVALID; }
}

/* debug / tuning: conv_tc2 pipeline profile (cycles per role, summed over CTAs; enabled by AID_TC_DEBUG bit 2048), read and cleared */
int aid_debug_tc2_profile(uint64_t* out16) {
    if (!out16) return AID_ERR_INVALID;
    try { tc2_read_profile(reinterpret_cast<unsigned long long*>(out16)); return AID_OK; } catch (const CudaError&) { return AID_ERR_CUDA; }
}

int aid_op_groupnorm_act(const float* x_dev, const float* gamma_dev, const float* affine_dev, int B, int C, int F, int T, int gelu,
                         float* out_dev, double* stats_scratch_dev, void* stream) {
    if (!x_dev || !gamma_dev || !out_dev || !stats_scratch_dev || C % 8 != 0) return AID_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    TV x = make_tv(const_cast<float*>(x_dev), B, C, F, T), out = make_tv(out_dev, B, C, F, T);
    if (cudaMemsetAsync(stats_scratch_dev, 0, (size_t)B * 16 * sizeof(double), s) != cudaSuccess) return AID_ERR_CUDA;
    launch_group_stats(x, stats_scratch_dev, s);
    launch_gn_act(x, stats_scratch_dev, (long long)(C / 8) * F * T, gamma_dev, affine_dev, 0, gelu != 0, out, s);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_op_resample(const float* x_dev, int B, int C, int F, int T, int up, float* out_dev, void* stream) {
    if (!x_dev || !out_dev || T < 4 || (T & 1)) return AID_ERR_INVALID;
    TV x = make_tv(const_cast<float*>(x_dev), B, C, F, T);
    TV out = make_tv(out_dev, B, C, F, up ? 2 * T : T / 2);
    if (up) launch_resample_up(x, out, (cudaStream_t)stream); else launch_resample_down(x, out, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_op_attention(const float* h_dev, const float* qk_dev, int B, int heads, int F, int T, float* out_dev, void* stream) {
    return aid_op_attention_mode(h_dev, qk_dev, B, heads, F, T, out_dev, 0, stream);
}

int aid_op_attention_mode(const float* h_dev, const float* qk_dev, int B, int heads, int F, int T, float* out_dev, int tensor_core, void* stream) {
    if (!h_dev || !qk_dev || !out_dev) return AID_ERR_INVALID;
    try {
        TV h = make_tv(const_cast<float*>(h_dev), B, heads, F, T), out = make_tv(out_dev, B, heads, F, T);
        if (tensor_core) {
            if (!attention_tc_supported(F, T)) return AID_ERR_INVALID;
            launch_attention_tc(h, qk_dev, out, (cudaStream_t)stream);
        } else launch_attention(h, qk_dev, out, (cudaStream_t)stream);
        return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
    } catch (...) { return AID_ERR_INVALID; }
}

/* ---- backward single-operator entry points (unit parity of the VJP pieces) ---------------------------------------------- */
int aid_op_resample_adj(const float* gy_dev, int B, int C, int F, int T, int up, float* gx_dev, void* stream) {
    if (!gy_dev || !gx_dev || T < 4 || (T & 1)) return AID_ERR_INVALID;
    TV gx = make_tv(gx_dev, B, C, F, T), gy = make_tv(const_cast<float*>(gy_dev), B, C, F, up ? 2 * T : T / 2);
    if (up) launch_resample_up_adj(gy, gx, 0.f, (cudaStream_t)stream); else launch_resample_down_adj(gy, gx, 0.f, (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_op_groupnorm_act_bwd(const float* g_dev, const float* x_dev, const float* gamma_dev, const float* affine_dev, int B, int C, int F, int T,
                             int gelu, float* gx_dev, double* scratch_dev /*[B][8][3]*/, void* stream) {
    if (!g_dev || !x_dev || !gamma_dev || !gx_dev || !scratch_dev || C % 8 != 0) return AID_ERR_INVALID;
    try {
        cudaStream_t s = (cudaStream_t)stream;
        TV x = make_tv(const_cast<float*>(x_dev), B, C, F, T), g = make_tv(const_cast<float*>(g_dev), B, C, F, T), gx = make_tv(gx_dev, B, C, F, T);
        AID_CUDA_CHECK(cudaMemsetAsync(scratch_dev, 0, (size_t)B * 24 * sizeof(double), s));
        launch_group_stats(x, scratch_dev, s);
        launch_gn_bwd(g, x, scratch_dev, (long long)(C / 8) * F * T, gamma_dev, affine_dev, 0, gelu != 0, scratch_dev + (size_t)B * 16, TV(), 0.f, gx, s);
        return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
    } catch (const CudaError&) { return AID_ERR_CUDA; }
}

int aid_op_attention_bwd(const float* h_dev, const float* qk_dev, const float* go_dev, int B, int heads, int F, int T, float* gh_dev,
                         float* gqk_dev, void* scratch_dev, size_t scratch_bytes, void* stream) {
    if (!h_dev || !qk_dev || !go_dev || !gh_dev || !gqk_dev || !scratch_dev) return AID_ERR_INVALID;
    try {
        Ctx c; c.B = B; c.s = (cudaStream_t)stream;
        c.ar.reset((char*)scratch_dev, scratch_bytes, false);
        TV h = make_tv(const_cast<float*>(h_dev), B, heads, F, T), qk = make_tv(const_cast<float*>(qk_dev), B, 2 * heads * F, 1, T);
        TV go = make_tv(const_cast<float*>(go_dev), B, heads, F, T), gh = make_tv(gh_dev, B, heads, F, T), gqk = make_tv(gqk_dev, B, 2 * heads * F, 1, T);
        attention_core_bwd(c, h, qk, go, gh, gqk);
        return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
    } catch (const std::exception&) { return AID_ERR_WORKSPACE; } catch (const CudaError&) { return AID_ERR_CUDA; }
}

/* transposed (data-gradient) convolution: gx[B,Cin,F,T] = conv^T(gy[B,Cout,F,T]; w[Cout,Cin,KF,KT]), the forward kernels with
 * tap-mirrored, channel-swapped weights */
int aid_op_conv2d_bwd_input(const float* gy_dev, const float* w_dev, int B, int Cin, int Cout, int F, int T, int KF, int KT, int dil,
                            float* gx_dev, void* stream) {
    if (!gy_dev || !w_dev || !gx_dev) return AID_ERR_INVALID;
    try {
        cudaStream_t s = (cudaStream_t)stream;
        const size_t e = (size_t)Cout * Cin * KF * KT;
        float *wp = nullptr, *wpT = nullptr;
        AID_CUDA_CHECK(cudaMalloc(&wp, e * sizeof(float))); AID_CUDA_CHECK(cudaMalloc(&wpT, e * sizeof(float)));
        pack_conv_weight_kernel<<<(int)std::min<size_t>(4096, (e + 255) / 256), 256, 0, s>>>(w_dev, wp, Cout, Cin, KF * KT);
        launch_pack_conv_weight_T(wp, wpT, Cout, Cin, KF, KT, s);
        TV gy = make_tv(const_cast<float*>(gy_dev), B, Cout, F, T), gx = make_tv(gx_dev, B, Cin, F, T);
        ConvEpilogue ep;
        if (!launch_conv_thin(gy, wpT, KF, KT, dil, gx, ep, s)) launch_conv_simt(gy, wpT, KF, KT, dil, gx, ep, s);
        AID_CUDA_CHECK(cudaGetLastError());
        AID_CUDA_CHECK(cudaStreamSynchronize(s));
        cudaFree(wp); cudaFree(wpT);
        return AID_OK;
    } catch (const CudaError&) { return AID_ERR_CUDA; }
}

/* adjoints of aid_cqt_fwd / aid_cqt_bwd (same coefficient layout and workspace) */
int aid_cqt_fwd_vjp(aid_handle* h, const float* gcoef_dev, float* gx_dev, int B, void* ws, size_t ws_bytes, void* stream) {
    if (!h || !gcoef_dev || !gx_dev || !ws) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net; cudaStream_t s = (cudaStream_t)stream;
        if (!n.finalized) throw std::runtime_error("aid_cqt_fwd_vjp before aid_finalize");
        ensure_vjp(n);
        float2 *spec, *tmp, *fscr, *Y; cqt_ws_split(n, B, ws, ws_bytes, &spec, &tmp, &fscr, &Y);
        const int L = n.cfg.audio_len, bins = n.cfg.bins_per_oct;
        int64_t off = 0;
        for (int o = 0; o < n.cfg.num_octs; ++o) {
            TV C = make_tv(const_cast<float*>(gcoef_dev) + off, B, 2, bins, n.tabs.M[o]);
            launch_cqt_synth_oct(n.tabs_adjA, n.fft, o, C, Y, s);
            off += (int64_t)B * 2 * bins * n.tabs.M[o];
        }
        launch_cqt_gather_adj(n.tabs_adjA, B, 0, n.cfg.num_octs - 1, Y, spec, false, s);
        launch_fft_big(n.fft, B, true, nullptr, 0, 0.f, spec, tmp, fscr, nullptr, gx_dev, L, 1.f, nullptr, 0, 0.f, s);
        AID_CUDA_CHECK(cudaGetLastError());
    });
}

int aid_cqt_bwd_vjp(aid_handle* h, const float* gx_dev, float* gcoef_dev, int B, void* ws, size_t ws_bytes, void* stream) {
    if (!h || !gcoef_dev || !gx_dev || !ws) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net; cudaStream_t s = (cudaStream_t)stream;
        if (!n.finalized) throw std::runtime_error("aid_cqt_bwd_vjp before aid_finalize");
        ensure_vjp(n);
        float2 *spec, *tmp, *fscr, *Y; cqt_ws_split(n, B, ws, ws_bytes, &spec, &tmp, &fscr, &Y);
        const int L = n.cfg.audio_len, bins = n.cfg.bins_per_oct;
        launch_fft_big(n.fft, B, false, gx_dev, L, 1.f, nullptr, tmp, fscr, spec, nullptr, 0, 0.f, nullptr, 0, 0.f, s);
        launch_spec_synth_adj(B, L, spec, 1.f / (float)L, s);
        int64_t off = 0;
        for (int o = 0; o < n.cfg.num_octs; ++o) {
            TV C = make_tv(gcoef_dev + off, B, 2, bins, n.tabs.M[o]);
            launch_cqt_analysis_oct(n.tabs_adjS, n.fft, o, spec, C, s);
            off += (int64_t)B * 2 * bins * n.tabs.M[o];
        }
        AID_CUDA_CHECK(cudaGetLastError());
    });
}

int aid_op_embedding(aid_handle* h, const float* c_noise_dev, int n_sigma, float* emb_dev, void* stream) {
    if (!h || !c_noise_dev || !emb_dev || !h->net.finalized) return AID_ERR_INVALID;
    Net& n = h->net;
    launch_embedding(c_noise_dev, n_sigma, n.d_emb[0], n.d_emb[1], n.d_emb[2], n.d_emb[3], n.d_emb[4], n.d_emb[5], n.d_emb[6], emb_dev,
                     (cudaStream_t)stream);
    return cudaGetLastError() == cudaSuccess ? AID_OK : AID_ERR_CUDA;
}

int aid_profile(aid_handle* h, int enable) {
    if (!h) return AID_ERR_INVALID;
    for (auto& r : h->net.prof_recs) { h->net.prof_pool.push_back(r.e0); h->net.prof_pool.push_back(r.e1); }
    h->net.prof_recs.clear();
    h->net.prof = enable != 0;
    return AID_OK;
}

int aid_profile_read(aid_handle* h, int kind, uint64_t* launches, double* ms, double* flops, double* bytes) {
    if (!h) return AID_ERR_INVALID;
    return guarded(h, [&] {
        uint64_t nl = 0; double t = 0, f = 0, b = 0;
        for (auto& r : h->net.prof_recs) {
            if (r.kind != kind) continue;
            AID_CUDA_CHECK(cudaEventSynchronize(r.e1));
            float e = 0.f;
            AID_CUDA_CHECK(cudaEventElapsedTime(&e, r.e0, r.e1));
            ++nl; t += e; f += r.flops; b += r.bytes;
        }
        if (launches) *launches = nl;
        if (ms) *ms = t;
        if (flops) *flops = f;
        if (bytes) *bytes = b;
    });
}

/* debug: conv_mode 2 converts its operands to fp16 with saturation (activations x16, weights x1024).  enable != 0 resets and
 * starts counting the activation values clamped by the operand passes of the following forwards; the counters are read
 * (after a device synchronisation) whenever a pointer is given.  weight_count: weight values clamped at aid_finalize. */
int aid_debug_saturation(aid_handle* h, int enable, uint64_t* act_count, uint64_t* weight_count) {
    if (!h) return AID_ERR_INVALID;
    return guarded(h, [&] {
        Net& n = h->net;
        if (!n.finalized) throw std::runtime_error("aid_debug_saturation before aid_finalize");
        if (weight_count) *weight_count = n.weight_sat;
        if (!n.d_sat) { if (act_count) *act_count = 0; n.count_sat = false; return; }   // conv_mode 0 / 1: nothing saturates
        DeviceGuard guard(n.device);
        if (act_count) {
            unsigned long long v = 0;
            AID_CUDA_CHECK(cudaDeviceSynchronize());
            AID_CUDA_CHECK(cudaMemcpy(&v, n.d_sat, sizeof v, cudaMemcpyDeviceToHost));
            *act_count = v;
        }
        if (enable && !n.count_sat) { AID_CUDA_CHECK(cudaDeviceSynchronize()); AID_CUDA_CHECK(cudaMemset(n.d_sat, 0, sizeof(unsigned long long))); }
        n.count_sat = enable != 0;
    });
}

int aid_debug_fusion(aid_handle* h, int init_blocks, int dilated_layers, int out_blocks, int upsampling) {
    if (!h) return AID_ERR_INVALID;
    if (init_blocks >= 0) h->net.fuse_init = init_blocks != 0;
    if (dilated_layers >= 0) h->net.fuse_comb = dilated_layers != 0;
    if (out_blocks >= 0) h->net.fuse_out = out_blocks != 0;
    if (upsampling >= 0) h->net.fuse_up = upsampling != 0;
    return AID_OK;
}

int aid_debug_probe(aid_handle* h, const char* name, float* dst_dev) {
    if (!h || !name) return AID_ERR_INVALID;
    if (dst_dev) h->net.probes[name] = dst_dev; else h->net.probes.erase(name);
    return AID_OK;
}

uint64_t aid_launch_count(void) { return aid::g_launch_count; }

}  // extern "C"
