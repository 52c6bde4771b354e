"""ORACLE (test infrastructure, not product): CPU restatement of `cqt_nsgt_pytorch.CQT_nsgt`.

PARITY UNPINNED.  The reference imports `cqt_nsgt_pytorch` (unet.py:9, pinned only as
`cqt_nsgt_pytorch-0.0.8` by notebooks/demo_inpainting_spectrogram.ipynb:149,153) but does not vendor
it, it is not installed in this image and there is no network.  No reference test pins its values.
This file restates the published algorithm (non-stationary Gabor transform, Velasco / Holighaus /
Doerfler / Grill 2011, "oct" mode of the upstream package) against the contract that IS readable in
the reference:

  ctor      unet.py:615-620   CQT_nsgt(num_octs, bins_per_oct, mode="oct", window=("kaiser", beta)|str,
                              fs=, audio_len=, dtype=, device=)
  fwd       unet.py:743,750-753   real [B,1,L] -> list of num_octs complex [B,1,bins,T_o], ascending
                              frequency, T_o doubling per octave (unet.py:769-774,786 needs that)
  bwd       unet.py:841-843   list -> real [B,1,>=L]
  apply_hpf_DC  sampler.py:63,123, edm.py:184   [B,L] -> [B,L], removes the DC/Nyquist bands

Only tests/, bench.py's cpu_baseline / reference arm and __graft_entry__.smoke() may import this.

Definition used here (everything in FFT-bin units, bin = f * L / fs):
  fmax = fs/2 - 1e-6, fmin = fmax / 2**num_octs, K = num_octs*bins bands, f_k = fmin * p**k,
  p = 2**(num_octs/(K-1)), Q = sqrt(p)/(p-1)/2.
  centre c_k = round(f_k) except the top band, centred halfway between f_{K-2} and Nyquist.
  width  Lg_k = round(f_{k+1} - f_{k-1}); first and last band round(f/Q); DC band round(2 f_0);
         Nyquist band max(4, .);  all clipped to >= 4.
  window w_k(m), m in [-floor(Lg/2), ceil(Lg/2)):  Kaiser(beta) (or Hann) sampled on the circle of
         length Lg with its peak at m = 0.
  "oct": every band of octave o gets M_o = next_pow2(max Lg in that octave) coefficients.
  analysis  coef_k = IFFT_{M_o}( fold( X[(c_k + m) mod L] * w_k(m) ) ),  X = FFT_L(x)
  synthesis fr[(c_k + m) mod L] += FFT_{M_o}(coef_k)[m mod M_o] * M_o * w_k(m) / D,  D = sum_j M_j w_j^2
            (sum over all bands incl. DC, Nyquist and the mirrored negative-frequency bands),
            x = irFFT_L(fr[:L/2+1])
  H_lpf = (M w^2 / D) of the DC band (+ Nyquist band), H_hpf = 1 - H_lpf.
"""
import math

import numpy as np
import torch


def _next_pow2(v: int) -> int:
    return 1 << max(0, int(v) - 1).bit_length()


def _window(Lg: int, window) -> np.ndarray:
    """w(m) stored in "peak at index 0" order: index j>=0 holds m=j for j<ceil(Lg/2), index Lg-j holds m=-j."""
    n = np.arange(Lg, dtype=np.float64)
    d = np.minimum(n, Lg - n)  # circular distance from the peak
    if isinstance(window, (tuple, list)) and window[0] == "kaiser":
        beta = float(window[1])
        r = 2.0 * d / Lg
        return np.i0(beta * np.sqrt(np.clip(1.0 - r * r, 0.0, 1.0))) / np.i0(beta)
    if window == "hann":
        return 0.5 * (1.0 + np.cos(2.0 * np.pi * d / Lg))
    raise ValueError(f"unsupported window {window!r}")


class CQTPlan:
    """Pure-numpy band plan shared by the oracle and (as tables) by the CUDA path's host code.

    The CUDA host code builds its own tables from the same definition (it must not import oracle/);
    tests compare the two plans entry by entry.
    """

    def __init__(self, numocts, binsoct, window, fs, audio_len, min_win=4):
        self.numocts, self.binsoct, self.fs, self.Ls = int(numocts), int(binsoct), float(fs), int(audio_len)
        L = self.Ls
        K = self.numocts * self.binsoct
        fmax = self.fs / 2.0 - 1e-6
        fmin = fmax / (2 ** self.numocts)
        p = 2.0 ** ((math.log2(fmax) - math.log2(fmin)) / (K - 1))
        q = math.sqrt(p) / (p - 1.0) / 2.0
        f = fmin * p ** np.arange(K, dtype=np.float64)
        nf = self.fs / 2.0
        frqs = np.concatenate(((0.0,), f, (nf,)))
        fbas = np.concatenate((frqs, self.fs - frqs[-2:0:-1])) * (L / self.fs)  # 2K+2 entries, in bins
        nb = 2 * K + 2
        M = np.zeros(nb, dtype=np.int64)
        M[0] = np.round(2.0 * fbas[1])
        M[1] = np.round(fbas[1] / q)
        for k in range(2, K):
            M[k] = np.round(fbas[k + 1] - fbas[k - 1])
        M[K] = np.round(fbas[K] / q)
        M[K + 1] = np.round(fbas[K + 2] - fbas[K])
        M[K + 2:] = M[K:0:-1]
        M = np.clip(M, min_win, None)
        self.Lg = M.copy()  # window lengths
        fb = fbas.copy()
        fb[K] = 0.5 * (fb[K - 1] + fb[K + 1])
        fb[K + 2] = L - fb[K]
        self.centre = np.round(fb).astype(np.int64)  # rfbas
        # "oct": coefficients per band = next pow2 of the widest window in the octave
        self.M = self.Lg.copy()
        self.size_per_oct = []
        idx = 1
        for _ in range(self.numocts):
            v = _next_pow2(int(self.Lg[idx:idx + self.binsoct].max()))
            self.size_per_oct.append(v)
            self.M[idx:idx + self.binsoct] = v
            self.M[nb - idx - self.binsoct + 1: nb - idx + 1] = v  # mirrored bands
            idx += self.binsoct
        self.K = K
        self.g = [_window(int(lg), window) for lg in self.Lg]
        # frame-operator diagonal D over the whole circle (all 2K+2 bands)
        D = np.zeros(L, dtype=np.float64)
        for k in range(nb):
            D[self.win_range(k)] += np.fft.fftshift(self.g[k]) ** 2 * self.M[k]
        self.D = D
        self.gd = [self.g[k] / np.fft.ifftshift(D[self.win_range(k)]) for k in range(nb)]
        # H_lpf: what the dropped DC and Nyquist bands would have reconstructed
        H = np.zeros(L, dtype=np.float64)
        for k in (0, K + 1):
            H[self.win_range(k)] += np.fft.fftshift(self.g[k] * self.gd[k]) * self.M[k]
        self.Hlpf = H
        self.Hhpf = 1.0 - H

    def win_range(self, k):
        lg = int(self.Lg[k])
        return (np.arange(-(lg // 2), lg - lg // 2, dtype=np.int64) + int(self.centre[k])) % self.Ls


class CQT_nsgt:
    """Drop-in for `cqt_nsgt_pytorch.CQT_nsgt` restricted to what the reference calls (mode="oct")."""

    def __init__(self, numocts, binsoct, mode="oct", window="hann", flex_Q=None, fs=44100, audio_len=44100,
                 device="cpu", dtype=torch.float32):
        if mode != "oct":
            raise NotImplementedError("the reference hot path only uses mode='oct' (unet.py:620)")
        self.numocts, self.binsoct, self.mode, self.fs, self.Ls = numocts, binsoct, mode, fs, audio_len
        self.device, self.dtype = torch.device(device), dtype
        self.plan = p = CQTPlan(numocts, binsoct, window, fs, audio_len)
        self.size_per_oct = list(p.size_per_oct)
        cdt = torch.complex64 if dtype == torch.float32 else torch.complex128
        self._oct = []
        for o in range(numocts):
            ks = range(1 + o * binsoct, 1 + (o + 1) * binsoct)
            Lmax = int(max(p.Lg[k] for k in ks))
            Mo = p.size_per_oct[o]
            idx = np.zeros((binsoct, Lmax), dtype=np.int64)
            win = np.zeros((binsoct, Lmax), dtype=np.float64)
            dual = np.zeros((binsoct, Lmax), dtype=np.float64)
            # column j of the (band, Lmax) table is frequency offset m = j - Lmax//2
            for r, k in enumerate(ks):
                lg = int(p.Lg[k])
                m = np.arange(-(lg // 2), lg - lg // 2)
                cols = m + Lmax // 2
                idx[r] = (p.centre[k] + np.arange(-(Lmax // 2), Lmax - Lmax // 2)) % p.Ls
                win[r, cols] = np.fft.fftshift(p.g[k])
                dual[r, cols] = np.fft.fftshift(p.gd[k]) * Mo  # synthesis gain M_o folded in
            self._oct.append(dict(
                idx=torch.from_numpy(idx).to(self.device), Lmax=Lmax, M=Mo,
                win=torch.from_numpy(win).to(dtype).to(self.device),
                dual=torch.from_numpy(dual).to(dtype).to(self.device)))
        self.Hhpf = torch.from_numpy(p.Hhpf).to(dtype).to(self.device)
        self._cdt = cdt

    # ---- analysis --------------------------------------------------------------------------------
    def fwd(self, x):
        """x real [B,C,L] -> list (ascending frequency) of complex [B,C,bins,M_o]."""
        assert x.shape[-1] == self.Ls
        ft = torch.fft.fft(x)
        out = []
        for o in self._oct:
            Lmax, M = o["Lmax"], o["M"]
            t = ft[..., o["idx"]] * o["win"]  # [B,C,bins,Lmax], column j <-> m = j - Lmax//2
            c = torch.zeros(*t.shape[:-1], M, dtype=t.dtype, device=t.device)
            c[..., :Lmax - Lmax // 2] = t[..., Lmax // 2:]  # m >= 0
            if Lmax // 2 > 0:
                c[..., M - Lmax // 2:] = t[..., :Lmax // 2]  # m < 0
            out.append(torch.fft.ifft(c))
        return out

    # ---- synthesis -------------------------------------------------------------------------------
    def bwd(self, coefs):
        """list of complex [B,C,bins,M_o] -> real [B,C,L]."""
        B, C = coefs[0].shape[:2]
        fr = torch.zeros(B, C, self.Ls, dtype=coefs[0].dtype, device=coefs[0].device)
        for o, c in zip(self._oct, coefs):
            Lmax, M = o["Lmax"], o["M"]
            assert c.shape[-1] == M and c.shape[-2] == self.binsoct
            t = torch.fft.fft(c)
            u = torch.cat((t[..., M - Lmax // 2:], t[..., :Lmax - Lmax // 2]), dim=-1) if Lmax // 2 > 0 \
                else t[..., :Lmax]
            u = u * o["dual"]
            fr.index_add_(-1, o["idx"].reshape(-1), u.reshape(B, C, -1))
        return torch.fft.irfft(fr[..., : self.Ls // 2 + 1], n=self.Ls)

    def apply_hpf_DC(self, x):
        Lin = x.shape[-1]
        if Lin > self.Ls:
            raise ValueError("input longer than audio_len")
        if Lin < self.Ls:
            x = torch.nn.functional.pad(x, (0, self.Ls - Lin))
        out = torch.fft.ifft(torch.fft.fft(x) * self.Hhpf).real
        return out[..., :Lin]


def install_as_cqt_nsgt_pytorch():
    """Register this restatement under the module name the reference imports (unet.py:9)."""
    import sys
    import types
    mod = types.ModuleType("cqt_nsgt_pytorch")
    mod.CQT_nsgt = CQT_nsgt
    mod.__oracle__ = True
    sys.modules["cqt_nsgt_pytorch"] = mod
    return mod
