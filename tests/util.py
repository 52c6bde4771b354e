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
