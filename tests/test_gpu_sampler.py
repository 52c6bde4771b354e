"""GPU parity of the sampler loop (fused CUDA steps + denoiser) against the oracle and the reference's golden run."""
import os

import numpy as np
import pytest
import torch

from util import rel_l2, seeded, make_oracle
from test_host import _tester_args

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden.npz")


@pytest.fixture(scope="module")
def setup(aid, cuda):
    cfg = aid.small_test(16384)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    return cfg, sd, net


def test_edm_step_kernels(aid, cuda):
    from aid_b200 import _lib
    L = _lib.lib()
    n, Ln = 3 * 5000, 5000
    x, xh, y, d0, xb = (seeded((3, Ln), s).to(cuda) for s in range(5))
    m = (torch.rand(Ln, generator=torch.Generator().manual_seed(9)) > 0.3).float().to(cuda)
    sigma, h = 0.37, -0.11
    d_out, x_out = torch.empty_like(x), torch.empty_like(x)
    _lib.check(L.aid_edm_step(_lib.ptr(x), _lib.ptr(xh), _lib.ptr(y), _lib.ptr(m), Ln, n, sigma, h, 0, None, None, _lib.ptr(d_out), _lib.ptr(x_out), None))
    proj = m * y + (1 - m) * xh
    d = (x - proj) / sigma
    assert rel_l2(d_out, d) < 1e-6 and rel_l2(x_out, x + h * d) < 1e-6
    _lib.check(L.aid_edm_step(_lib.ptr(x), _lib.ptr(xh), None, None, 0, n, sigma, h, 1, _lib.ptr(d0), _lib.ptr(xb), None, _lib.ptr(x_out), None))
    d = (x - xh) / sigma
    assert rel_l2(x_out, xb + h * (0.5 * d0 + 0.5 * d)) < 1e-6
    e = seeded((3, Ln), 8).to(cuda)
    x2 = x.clone()
    _lib.check(L.aid_edm_add_noise(_lib.ptr(x2), _lib.ptr(e), 0.25, n, None))
    assert rel_l2(x2, x + 0.25 * e) < 1e-6


def test_sampler_matches_reference_golden(aid, cuda, setup):
    """Same seeds as tests/golden/make_golden.py: the reference Sampler + EDM + unet.py produced these on the CPU.
    Six Heun steps chain 11 denoiser calls, each within 1e-4 of the oracle; the trajectory is held to 1e-3."""
    cfg, sd, net = setup
    g = np.load(GOLD)
    args = _tester_args(aid, T=6)
    s = aid.Sampler(net, aid.EDM(args), args)
    torch.manual_seed(42)
    xu = s.predict_unconditional((2, cfg.audio_len), cuda)
    assert xu.is_cuda and rel_l2(xu, torch.from_numpy(g["small_sample_uncond_T6"])) < 1e-3
    y = seeded((2, cfg.audio_len), 7, 0.063)
    mask = torch.ones(1, cfg.audio_len)
    mask[..., cfg.audio_len // 2 - 750: cfg.audio_len // 2 + 750] = 0
    torch.manual_seed(43)
    xi = s.predict_inpainting((y * mask).to(cuda), mask.to(cuda))
    assert rel_l2(xi, torch.from_numpy(g["small_sample_inpaint_T6"])) < 1e-3
    keep = mask[0].bool()
    keep[cfg.audio_len // 2 - 850: cfg.audio_len // 2 + 850] = False
    assert torch.allclose(xi[:, keep.to(cuda)], (y * mask)[:, keep].to(cuda), atol=1e-6)


def test_sampler_fused_equals_generic_path(aid, cuda, setup):
    """The fused CUDA step kernels against the reference-ordered torch ops driving the same CUDA denoiser."""
    cfg, sd, net = setup
    args = _tester_args(aid, T=4)
    y = seeded((2, cfg.audio_len), 3, 0.063).to(cuda)
    mask = torch.ones(1, cfg.audio_len, device=cuda)
    mask[..., 7000:9000] = 0
    s = aid.Sampler(net, aid.EDM(args), args)
    torch.manual_seed(1)
    fused = s.predict_inpainting(y * mask, mask)

    class Plain(torch.nn.Module):  # hides denoise_fused -> Sampler takes the torch-op path
        CQTransform = net.CQTransform

        def forward(self, x, c):
            return net(x, c)

    s2 = aid.Sampler(Plain(), aid.EDM(args), args)
    torch.manual_seed(1)
    plain = s2.predict_inpainting(y * mask, mask)
    assert rel_l2(fused, plain) < 1e-4


def test_sampler_single_fp16_path_vs_reference_golden(aid, cuda):
    """conv_mode 2 (one fp16 MMA per tap) through the whole sampler: every denoiser call is within 1e-3 of the reference
    (tests/test_gpu_tc.py); six Heun steps chain 11 calls, the trajectory is held to 3e-3 against the reference's own run."""
    cfg = aid.small_test(16384, conv_mode=2)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1234))
    g = np.load(GOLD)
    args = _tester_args(aid, T=6)
    s = aid.Sampler(net, aid.EDM(args), args)
    torch.manual_seed(42)
    xu = s.predict_unconditional((2, cfg.audio_len), cuda)
    eu = rel_l2(xu, torch.from_numpy(g["small_sample_uncond_T6"]))
    y = seeded((2, cfg.audio_len), 7, 0.063)
    mask = torch.ones(1, cfg.audio_len)
    mask[..., cfg.audio_len // 2 - 750: cfg.audio_len // 2 + 750] = 0
    torch.manual_seed(43)
    xi = s.predict_inpainting((y * mask).to(cuda), mask.to(cuda))
    ei = rel_l2(xi, torch.from_numpy(g["small_sample_inpaint_T6"]))
    print(f"conv_mode 2 sampler trajectories vs reference golden: unconditional {eu:.3e}, inpainting {ei:.3e}")
    assert eu < 3e-3 and ei < 3e-3


GAPS_MS = {25: 551, 50: 1102, 100: 2205, 300: 6615, 371: 8180, 743: 16383, 1486: 32766, 1500: 33075}


@pytest.mark.parametrize("mode", [0, 2])
def test_inpainting_gap_sweep_vs_oracle(aid, cuda, mode):
    """BASELINE config 5 as a correctness sweep: the reference's gap lengths (inpainting_tester.yaml:71,75;
    tester_inpainting.py:355-357) in samples at 22.05 kHz, centred like prepare_mask, three Heun steps (5 denoiser calls) on a
    65536-sample clip.  Each gap is compared with the oracle sampler driving the oracle network on the same noise; outside
    the gap and its 50-sample ramps the known samples must come back exactly."""
    import unet_oracle
    L = 65536
    cfg = aid.small_test(L, conv_mode=mode)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    oracle = make_oracle(cfg, sd)
    args = _tester_args(aid, T=3)
    s = aid.Sampler(net, aid.EDM(args), args)
    y = seeded((1, L), 7, 0.063)
    tol = 1e-3 if mode == 0 else 3e-3           # 5 chained calls; the per-call bars are 1e-4 / 1e-3

    def stream():
        while True:
            yield torch.randn(1, L)

    worst = 0.0
    for ms, gap in GAPS_MS.items():
        mask = torch.ones(1, L)
        a = L // 2 - gap // 2
        mask[..., a:a + gap] = 0
        torch.manual_seed(100 + ms)
        got = s.predict_inpainting((y * mask).to(cuda), mask.to(cuda)).cpu()
        torch.manual_seed(100 + ms)
        want = unet_oracle.sample_oracle(oracle, unet_oracle.EDMOracle(), (1, L), stream(), nb_steps=3, y=y * mask,
                                         mask_s=unet_oracle.smooth_mask(mask, 50))
        e = rel_l2(got, want)
        worst = max(worst, e)
        assert e < tol, (ms, gap, e)
        keep = torch.ones(L, dtype=torch.bool)
        keep[a - 50:a + gap + 50] = False
        assert torch.allclose(got[:, keep], y[:, keep], atol=1e-6), ms
        assert got[:, a:a + gap].abs().max() > 0
    print(f"conv_mode {mode}: worst gap-sweep rel-L2 vs oracle sampler {worst:.3e}")


@pytest.mark.parametrize("name", ["ragged", "multiple", "small_random"])
def test_spectral_mask_kernels_vs_reference_golden(aid, cuda, name):
    """csrc/stft.cu through aid_spectral_mask against the output of the reference's own Sampler.apply_spectral_mask
    (tests/golden/golden_spectral.npz): the degradation S(x) and the projection y + x - S(x).  fp32 FFTs: 1e-5."""
    from util import spectral_case
    x, mask, n_fft, hop = spectral_case(name)
    want = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_spectral.npz"))[name])
    args = _tester_args(aid)
    args["tester"]["spectrogram_inpainting"]["stft"].update({"n_fft": n_fft, "hop_length": hop, "win_length": n_fft})
    s = aid.Sampler(torch.nn.Identity(), aid.EDM(args), args)
    s.mask = mask.to(cuda)
    got = s.apply_spectral_mask(x.to(cuda))
    assert got.is_cuda and got.shape == x.shape
    assert rel_l2(got, want) < 1e-5
    y = seeded(tuple(x.shape), 12, 0.063)
    proj = s.apply_spectral_mask(x.to(cuda), y.to(cuda))
    assert rel_l2(proj, y + x - want) < 1e-5
    with pytest.raises(ValueError, match="spectral mask"):
        s.mask = mask[:, :-1].to(cuda)
        s.apply_spectral_mask(x.to(cuda))


def test_spectral_mask_full_size_vs_oracle(aid, cuda):
    """BASELINE-size clips (4 x 262144, the reference's 1024 / 256 STFT and its 2 s x 300-2000 Hz mask) against the oracle."""
    import unet_oracle
    from util import spectral_mask_rect
    L = 262144
    x = seeded((4, L), 3, 0.063)
    mask = spectral_mask_rect(L)
    args = _tester_args(aid)
    s = aid.Sampler(torch.nn.Identity(), aid.EDM(args), args)
    s.mask = mask.to(cuda)
    got = s.apply_spectral_mask(x.to(cuda))
    want = unet_oracle.spectral_mask(x, mask)
    assert rel_l2(got, want) < 1e-5
    assert rel_l2(x.to(cuda) - got, x - want) < 1e-4        # the removed band itself (small difference of large numbers)
    ones = torch.ones_like(mask).to(cuda)                    # an all-pass mask reconstructs the input
    s.mask = ones
    assert rel_l2(s.apply_spectral_mask(x.to(cuda)), x) < 1e-5


@pytest.mark.parametrize("mode", [0, 2])
def test_spectrogram_inpainting_sampler_vs_oracle(aid, cuda, mode):
    """predict_spectrogram_inpainting (sampler.py:348-364) on the CUDA path -- fused denoiser, STFT projection kernels, fused
    step -- against the oracle sampler driving the oracle network with the same noise: 3 Heun steps, 65536-sample clips."""
    import unet_oracle
    from util import spectral_mask_rect
    L = 65536
    cfg = aid.small_test(L, conv_mode=mode)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    oracle = make_oracle(cfg, sd)
    args = _tester_args(aid, T=3)
    s = aid.Sampler(net, aid.EDM(args), args)
    y = seeded((2, L), 7, 0.063)
    mask = spectral_mask_rect(L, gap_ms=500)
    y_masked = unet_oracle.spectral_mask(y, mask)

    def stream():
        while True:
            yield torch.randn(2, L)

    torch.manual_seed(77)
    got = s.predict_spectrogram_inpainting(y_masked.to(cuda), mask.to(cuda))
    torch.manual_seed(77)
    want = unet_oracle.sample_oracle(oracle, unet_oracle.EDMOracle(), (2, L), stream(), nb_steps=3, y=y_masked,
                                     project=unet_oracle.spectral_projection(y_masked, mask))
    e = rel_l2(got, want)
    print(f"conv_mode {mode}: spectrogram-inpainting trajectory vs oracle {e:.3e}")
    assert got.is_cuda and e < (1e-3 if mode == 0 else 3e-3)
