"""Shared helpers for the parity tests (oracle construction, error metrics)."""
import torch


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def oracle_cfg(cfg):
    return dict(num_octs=cfg.num_octs, bins_per_oct=cfg.bins_per_oct, sample_rate=cfg.sample_rate, audio_len=cfg.audio_len,
                window=cfg.window, beta=cfg.beta, Ns=cfg.Ns, num_dils=cfg.num_dils, attention_layers=cfg.attention_layers)


def make_oracle(cfg, sd):
    import unet_oracle
    return unet_oracle.UnetOracle(oracle_cfg(cfg), sd)


def seeded(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def spectral_mask_rect(L, sample_rate=22050, n_fft=1024, hop=256, gap_ms=2000, fmin=300, fmax=2000):
    """The rectangular [n_fft/2+1, frames] mask of tester_inpainting.py:256-296 (time gap centred, fmin..fmax zeroed)."""
    frames = 1 + (L + n_fft - L % n_fft) // hop
    A = torch.ones(n_fft // 2 + 1, frames)
    freqs = torch.fft.fftfreq(n_fft, d=1 / sample_rate)
    f0, f1 = int(torch.argmin(torch.abs(freqs - fmin))), int(torch.argmin(torch.abs(freqs - fmax)))
    gap = int(gap_ms * sample_rate / 1000)
    start = int(L // 2 - gap // 2) // hop
    A[f0:f1, start:start + gap // hop] = 0
    return A


SPECTRAL_CASES = {  # name: (B, L, n_fft, hop, mask kind) -- the cases of tests/golden/golden_spectral.npz
    "ragged": (2, 20000, 1024, 256, "rect"),
    "multiple": (1, 16384, 1024, 256, "rect"),
    "small_random": (3, 3000, 256, 64, "random"),
}


def spectral_case(name):
    """(x, mask, n_fft, hop) of a golden case, seeded exactly as tests/golden/make_golden_spectral.py did."""
    B, L, n_fft, hop, kind = SPECTRAL_CASES[name]
    x = seeded((B, L), 11, 0.063)
    if kind == "rect":
        mask = spectral_mask_rect(L, n_fft=n_fft, hop=hop, gap_ms=300)
    else:
        frames = 1 + (L + n_fft - L % n_fft) // hop
        mask = (torch.rand(n_fft // 2 + 1, frames, generator=torch.Generator().manual_seed(3)) > 0.4).float()
    return x, mask, n_fft, hop
