// Host-side band plan of the constant-Q NSGT ("oct" mode), double precision.
// Definition: DESIGN.md section "CQT" (the un-vendored cqt_nsgt_pytorch==0.0.8 called from unet.py:615-620).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace aid {

struct CqtPlanHost {
    int L = 0, K = 0, bins = 0, nocts = 0;
    std::vector<int> Lg_all, centre_all, M_all;   // all 2K+2 bands (DC, K bands, Nyquist, mirrored)
    std::vector<int> size_per_oct;                // M_o
    // tables for the K analysed bands (plan bands 1..K)
    std::vector<int> centre, Lg, woff, klo, khi;
    std::vector<float> win, dual, hhpf;

    static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

    static double window_at(int d, int lg, int kind, double beta) {
        // d = circular distance from the window peak; kind 0 = Hann, 1 = Kaiser(beta)
        if (kind == 1) {
            const double r = 2.0 * d / lg;
            double a = 1.0 - r * r; if (a < 0) a = 0;
            return std::cyl_bessel_i(0.0, beta * std::sqrt(a)) / std::cyl_bessel_i(0.0, beta);
        }
        return 0.5 * (1.0 + std::cos(2.0 * M_PI * d / lg));
    }

    void build(int nocts_, int bins_, double fs, int L_, int window_kind, double beta) {
        nocts = nocts_; bins = bins_; L = L_; K = nocts * bins;
        if (L < 4096 || (L & 1) != 0 || L > (1 << 21))
            throw std::invalid_argument("audio_len must be even and in [4096, 2097152]");
        if (nocts > 16) throw std::invalid_argument("num_octs > 16");
        const int min_win = 4;
        const double fmax = fs / 2.0 - 1e-6, fmin = fmax / std::pow(2.0, nocts);
        const double p = std::pow(2.0, (std::log2(fmax) - std::log2(fmin)) / (K - 1));
        const double q = std::sqrt(p) / (p - 1.0) / 2.0;
        const int nb = 2 * K + 2;
        std::vector<double> fbas(nb);
        fbas[0] = 0.0;
        for (int k = 0; k < K; ++k) fbas[1 + k] = fmin * std::pow(p, (double)k);
        fbas[K + 1] = fs / 2.0;
        for (int j = K + 2; j < nb; ++j) fbas[j] = fs - fbas[nb - j];  // mirror of band nb-j
        for (auto& v : fbas) v *= (double)L / fs;
        Lg_all.assign(nb, 0);
        auto rnd = [](double v) { return (long long)std::nearbyint(v); };  // round-half-even, like numpy
        Lg_all[0] = (int)rnd(2.0 * fbas[1]);
        Lg_all[1] = (int)rnd(fbas[1] / q);
        for (int k = 2; k < K; ++k) Lg_all[k] = (int)rnd(fbas[k + 1] - fbas[k - 1]);
        Lg_all[K] = (int)rnd(fbas[K] / q);
        Lg_all[K + 1] = (int)rnd(fbas[K + 2] - fbas[K]);
        for (int j = K + 2; j < nb; ++j) Lg_all[j] = Lg_all[nb - j];
        for (auto& v : Lg_all) if (v < min_win) v = min_win;
        std::vector<double> fb = fbas;
        fb[K] = 0.5 * (fb[K - 1] + fb[K + 1]);
        fb[K + 2] = (double)L - fb[K];
        centre_all.resize(nb);
        for (int j = 0; j < nb; ++j) centre_all[j] = (int)rnd(fb[j]);
        M_all = Lg_all;
        size_per_oct.clear();
        for (int o = 0; o < nocts; ++o) {
            int mx = 0;
            for (int k = 1 + o * bins; k < 1 + (o + 1) * bins; ++k) mx = Lg_all[k] > mx ? Lg_all[k] : mx;
            const int v = next_pow2(mx);
            if (v > 4096) throw std::invalid_argument("band transform longer than 4096 frames is not supported");
            size_per_oct.push_back(v);
            for (int k = 1 + o * bins; k < 1 + (o + 1) * bins; ++k) { M_all[k] = v; M_all[nb - k] = v; }
        }
        // the U-Net needs frame counts that double per octave (unet.py:769-774, 786)
        for (int o = 1; o < nocts; ++o)
            if (size_per_oct[o] != 2 * size_per_oct[o - 1])
                throw std::invalid_argument("CQT frame counts do not double per octave for this (fs, audio_len)");
        // windows (peak at index 0) for every band, frame-operator diagonal D
        std::vector<std::vector<double>> g(nb);
        for (int j = 0; j < nb; ++j) {
            const int lg = Lg_all[j];
            g[j].resize(lg);
            for (int n = 0; n < lg; ++n) g[j][n] = window_at(n < lg - n ? n : lg - n, lg, window_kind, beta);
        }
        auto bin_of = [&](int j, int m) { long long b = ((long long)centre_all[j] + m) % L; if (b < 0) b += L; return (int)b; };
        auto widx = [](int m, int lg) { return m >= 0 ? m : lg + m; };
        std::vector<double> D(L, 0.0);
        for (int j = 0; j < nb; ++j) {
            const int lg = Lg_all[j];
            for (int m = -(lg / 2); m < lg - lg / 2; ++m) { const double w = g[j][widx(m, lg)]; D[bin_of(j, m)] += w * w * M_all[j]; }
        }
        std::vector<double> H(L, 0.0);
        for (int j : {0, K + 1}) {
            const int lg = Lg_all[j];
            for (int m = -(lg / 2); m < lg - lg / 2; ++m) { const double w = g[j][widx(m, lg)]; const int b = bin_of(j, m); H[b] += w * w / D[b] * M_all[j]; }
        }
        hhpf.resize(L);
        for (int n = 0; n < L; ++n) hhpf[n] = (float)(1.0 - H[n]);
        centre.resize(K); Lg.resize(K); woff.resize(K);
        win.clear(); dual.clear();
        for (int k = 0; k < K; ++k) {
            const int j = k + 1, lg = Lg_all[j];
            centre[k] = centre_all[j]; Lg[k] = lg; woff[k] = (int)win.size();
            for (int n = 0; n < lg; ++n) {
                const int m = n < lg - lg / 2 ? n : n - lg;
                win.push_back((float)g[j][n]);
                dual.push_back((float)(g[j][n] / D[bin_of(j, m)] * M_all[j]));
            }
        }
        const int half = L / 2;
        klo.assign(half + 1, K); khi.assign(half + 1, -1);
        for (int k = 0; k < K; ++k) {
            const int lg = Lg[k];
            for (int m = -(lg / 2); m < lg - lg / 2; ++m) {
                const int n = centre[k] + m;  // positive-frequency bands never wrap below 0; bins above L/2 are dropped (irfft)
                if (n < 0 || n > half) continue;
                if (k < klo[n]) klo[n] = k;
                if (k > khi[n]) khi[n] = k;
            }
        }
    }
};

}  // namespace aid
