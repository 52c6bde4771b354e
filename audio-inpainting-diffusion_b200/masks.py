"""Masks that feed the sampler: the time-domain gap masks and the rectangular spectrogram mask the reference's tester builds
before it calls `Sampler.predict_inpainting` / `predict_spectrogram_inpainting` (testing/tester_inpainting.py:231-254 and
256-296).  Pure host code (index arithmetic on the config), same config keys, same results.

    mask  = prepare_mask(args, device)             # [1, audio_len], 0 inside the gap(s)
    smask = prepare_spectral_mask(args, device)    # [n_fft/2 + 1, frames], 0 inside the time-frequency rectangle
"""
import torch

from .config import cfg_get


def prepare_mask(args, device="cpu"):
    """tester_inpainting.py:231-254.  mask_mode "long": one gap of `long.gap_length` ms, centred unless `long.start_gap_idx`
    (ms) is given; "short": `short.num_gaps` gaps of `short.gap_length` ms at positions drawn with torch.randint (global
    CPU generator, as the reference does)."""
    L, fs = int(cfg_get(args, "exp.audio_len")), cfg_get(args, "exp.sample_rate")
    mask = torch.ones((1, L))
    mode = cfg_get(args, "tester.inpainting.mask_mode")
    if mode == "long":
        gap = int(cfg_get(args, "tester.inpainting.long.gap_length") * fs / 1000)
        start = cfg_get(args, "tester.inpainting.long.start_gap_idx")
        start = int(L // 2 - gap // 2) if start == "None" else int(start * fs / 1000)
        mask[..., start:start + gap] = 0
    elif mode == "short":
        n = int(cfg_get(args, "tester.inpainting.short.num_gaps"))
        gap = int(cfg_get(args, "tester.inpainting.short.gap_length") * fs / 1000)
        if cfg_get(args, "tester.inpainting.short.start_gap_idx") != "None":
            raise NotImplementedError                                       # tester_inpainting.py:252
        starts = torch.randint(0, L - gap, (n,))
        for i in range(n):
            mask[..., starts[i]:starts[i] + gap] = 0
    return mask.to(device)


def spectral_frames(audio_len, n_fft, hop):
    """Frames of torch.stft (centre = True) after the reference's zero padding to a multiple of n_fft (a whole n_fft when the
    length already is one): the second dimension of the spectrogram mask."""
    return 1 + (audio_len + n_fft - audio_len % n_fft) // hop


def prepare_spectral_mask(args, device="cpu"):
    """tester_inpainting.py:256-296: ones [n_fft/2+1, frames] with the bins between `min_masked_freq` and `max_masked_freq`
    zeroed over `time_mask_length` ms, centred unless `time_start_idx` (ms) is given."""
    si = "tester.spectrogram_inpainting."
    if cfg_get(args, si + "stft.window") != "hann":
        raise NotImplementedError("Only hann window is implemented for now")
    L, fs = int(cfg_get(args, "exp.audio_len")), cfg_get(args, "exp.sample_rate")
    n_fft, hop = int(cfg_get(args, si + "stft.n_fft")), int(cfg_get(args, si + "stft.hop_length"))
    A = torch.ones((n_fft // 2 + 1, spectral_frames(L, n_fft, hop)))
    freqs = torch.fft.fftfreq(n_fft, d=1 / fs)
    f0 = torch.argmin(torch.abs(freqs - cfg_get(args, si + "min_masked_freq")))
    f1 = torch.argmin(torch.abs(freqs - cfg_get(args, si + "max_masked_freq")))
    gap = int(cfg_get(args, si + "time_mask_length") * fs / 1000)
    start = cfg_get(args, si + "time_start_idx")
    start = int(L // 2 - gap // 2) // hop if start == "None" else int(start * fs / 1000) // hop
    A[f0:f1, start:start + gap // hop] = 0
    return A.to(device)
