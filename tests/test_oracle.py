"""CPU tests of the oracle itself: pinned against the golden vectors the reference's own code produced
(tests/golden/make_golden.py) and against the frame identities of the CQT definition."""
import json
import os

import numpy as np
import pytest
import torch

from util import rel_l2, seeded, make_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLD, "golden.npz")), json.load(open(os.path.join(GOLD, "golden_meta.json")))


@pytest.fixture(scope="module")
def small(aid):
    cfg = aid.small_test(16384)
    sd = aid.random_state_dict(cfg, seed=1234)
    return cfg, sd, make_oracle(cfg, sd)


def test_unet_oracle_matches_reference_golden(small, golden):
    cfg, sd, orc = small
    g, _ = golden
    x = seeded((2, cfg.audio_len), 0)
    for i, cn in enumerate([0.0613, -0.75, -2.3]):
        assert rel_l2(orc(x, torch.tensor([[cn]])), torch.from_numpy(g[f"small_fwd_{i}"])) < 1e-5
    x3 = seeded((3, cfg.audio_len), 3, 0.3)
    out = orc(x3, torch.tensor([[0.05], [-0.8], [-1.7]]))
    assert rel_l2(out, torch.from_numpy(g["small_fwd_persample"])) < 1e-5


def test_fixture_is_sensitive_to_every_branch(small):
    """Finding 6 of SURVEY.md: with default init the gated branches vanish; the test weights must not allow that."""
    cfg, sd, orc = small
    x = seeded((1, cfg.audio_len), 0)
    cn = torch.tensor([[-0.5]])
    pb = {}
    base = orc(x, cn, probe=pb)
    # (weight, block output it must visibly change).  Deep blocks move the waveform little with random weights,
    # which is why the GPU parity tests also compare these block outputs (test_forward_blockwise_vs_oracle).
    for key, where in (("downs.6.2.H.1.weight", "enc6"), ("ups.0.1.attn_block.qk.weight", "dec0"),
                       ("downs.4.2.attn_block.qk.weight", "enc4"), ("middle.0.1.gate2.bias", "mid"),
                       ("middle.0.1.attn_block.proj_in.weight", "mid"), ("ups.1.1.affine2.weight", "dec1"),
                       ("downs.3.1.weight", "enc4"), ("downs.0.0.norm.0.gamma", "enc0"), ("ups.3.1.res_conv.weight", "dec3")):
        sd2 = dict(sd)
        sd2[key] = sd[key] * 0.0 if "gamma" not in key else sd[key] * 1.5
        p2 = {}
        out = make_oracle(cfg, sd2)(x, cn, probe=p2)
        assert rel_l2(p2[where], pb[where]) > 2e-3, (key, where)  # 20x the 1e-4 block-wise GPU tolerance
        assert rel_l2(out, base) > 5e-5, key


def test_sampler_oracle_matches_reference_golden(small, golden):
    import unet_oracle
    cfg, sd, orc = small
    g, meta = golden
    edm = unet_oracle.EDMOracle()
    t = edm.create_schedule(35)
    assert torch.allclose(t, torch.tensor(meta["schedule_T35"]), rtol=0, atol=0)
    assert torch.allclose(edm.get_gamma(t), torch.tensor(meta["gamma_T35"]), rtol=0, atol=0)
    shape = (2, cfg.audio_len)

    def stream():
        while True:
            yield torch.randn(shape)

    torch.manual_seed(42)
    xu = unet_oracle.sample_oracle(orc, edm, shape, stream(), nb_steps=6, hpf=orc.CQTransform.apply_hpf_DC)
    assert rel_l2(xu, torch.from_numpy(g["small_sample_uncond_T6"])) < 1e-4
    y = seeded(shape, 7, 0.063)
    mask = torch.ones(1, cfg.audio_len)
    mask[..., cfg.audio_len // 2 - 750: cfg.audio_len // 2 + 750] = 0
    ms = unet_oracle.smooth_mask(mask.expand(2, -1), 50)
    torch.manual_seed(43)
    xi = unet_oracle.sample_oracle(orc, edm, shape, stream(), nb_steps=6, y=y * mask, mask_s=ms)
    assert rel_l2(xi, torch.from_numpy(g["small_sample_inpaint_T6"])) < 1e-4


def test_paper_network_oracle_matches_reference_golden(aid, golden):
    """BASELINE config 1 (186 M parameters, 1 x 65536) through EDM.denoiser."""
    import unet_oracle
    g, meta = golden
    cfg = aid.paper_22k(65536)
    assert [[k, list(s)] for k, s in aid.schema_from_lib(cfg)].sort() == meta["schema_paper"].sort()
    sd = aid.random_state_dict(cfg, seed=1234)
    orc = make_oracle(cfg, sd)
    edm = unet_oracle.EDMOracle()
    x = seeded((1, 65536), 0)
    assert rel_l2(edm.denoiser(x, orc, torch.tensor([1.0])), torch.from_numpy(g["paper_denoise_0"])) < 1e-5
    assert rel_l2(edm.denoiser(x * 0.05, orc, torch.tensor([0.05])), torch.from_numpy(g["paper_denoise_1"])) < 1e-5


@pytest.mark.parametrize("L,T0", [(16384, 256), (65536, 1024), (262144, 4096)])
def test_cqt_oracle_contract(L, T0):
    """unet.py:769-774 needs frame counts that double per octave; App. B: T0 = 4096 at L = 262144, 1024 at 65536."""
    import cqt_oracle
    c = cqt_oracle.CQT_nsgt(7, 64, "oct", ("kaiser", 1), fs=22050, audio_len=L)
    assert c.size_per_oct == [T0 >> (6 - o) for o in range(7)]
    if L > 65536:
        return
    x = seeded((2, 1, L), 1)
    X = c.fwd(x)
    assert [tuple(a.shape) for a in X] == [(2, 1, 64, T0 >> (6 - o)) for o in range(7)] and X[0].dtype == torch.complex64
    # linearity
    x2 = seeded((2, 1, L), 2)
    X12 = c.fwd(2.0 * x - 0.5 * x2)
    for a, b, d in zip(c.fwd(x), c.fwd(x2), X12):
        assert rel_l2(torch.view_as_real(2.0 * a - 0.5 * b), torch.view_as_real(d)) < 1e-5
    # frame identity on the half spectrum, below the Nyquist-straddling top band
    p = c.plan
    top = int(p.centre[p.K] - p.Lg[p.K] // 2)
    Y = torch.fft.rfft(c.bwd(X)[:, 0].double())[:, :top]
    Xf = torch.fft.rfft(x[:, 0].double())[:, :top]
    H = torch.from_numpy(p.Hhpf[:top])
    assert ((Y - Xf * H).norm() / (Xf * H).norm()).item() < 1e-5
    # apply_hpf_DC removes DC, keeps the mid band
    h = c.apply_hpf_DC(x[:, 0])
    assert abs(h.mean().item()) < 1e-6
    Hf, X0 = torch.fft.rfft(h.double()), torch.fft.rfft(x[:, 0].double())
    mid = slice(int(p.centre[2]), top)
    assert ((Hf[:, mid] - X0[:, mid]).norm() / X0[:, mid].norm()).item() < 1e-5


def test_resampler_oracle_facts():
    """down has DC gain 1 and maps an impulse at 8 to taps at outputs 2..5; up has DC gain 0.5."""
    import unet_oracle
    x = torch.ones(1, 1, 1, 32)
    assert torch.allclose(unet_oracle.down_t(x), torch.ones(1, 1, 1, 16), atol=1e-6)
    assert torch.allclose(unet_oracle.up_t(x), 0.5 * torch.ones(1, 1, 1, 64), atol=1e-6)
    imp = torch.zeros(1, 1, 1, 32)
    imp[..., 8] = 1
    d = unet_oracle.down_t(imp)[0, 0, 0]
    assert torch.allclose(d[2:6], torch.tensor([-0.01171875, 0.11328125, 0.43359375, -0.03515625]), atol=1e-7)


def _stft_kernel_definition(x, mask, n_fft, hop):
    """What csrc/stft.cu computes, stated with numpy FFTs (frame grid, reflect / zero-extension indices, mask on both halves of
    the full spectrum, synthesis window, gather overlap-add with the squared-window envelope)."""
    import numpy as np
    B, L = x.shape
    N = n_fft
    Lp = L + (N - L % N)
    n_frames = 1 + Lp // hop
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)
    xs = x.numpy().astype(np.float64)
    frames = np.zeros((B, n_frames, N))
    kk = np.minimum(np.arange(N), N - np.arange(N))
    for t in range(n_frames):
        i = t * hop - N // 2 + np.arange(N)
        i = np.where(i < 0, -i, np.where(i >= Lp, 2 * (Lp - 1) - i, i))
        v = np.where(i < L, xs[:, np.minimum(i, L - 1)], 0.0)
        X = np.fft.fft(w * v, axis=-1) * mask.numpy()[kk, t]
        frames[:, t] = w * np.fft.ifft(X, axis=-1).real
    out = np.zeros((B, L))
    for n in range(L):
        p = n + N // 2
        t_lo = (p - N) // hop + 1 if p >= N else 0
        t_hi = min(p // hop, n_frames - 1)
        ts = np.arange(t_lo, t_hi + 1)
        j = p - ts * hop
        out[:, n] = frames[:, ts, j].sum(-1) / (w[j] ** 2).sum()
    return torch.from_numpy(out).float()


@pytest.mark.parametrize("name", ["ragged", "multiple", "small_random"])
def test_spectral_mask_oracle_and_kernel_definition_match_reference_golden(name):
    """oracle.spectral_mask (sampler.py:271-290 restated with the same torch calls) is bit-identical to what the reference's own
    Sampler.apply_spectral_mask produced (tests/golden/make_golden_spectral.py); the formula the CUDA kernels implement agrees
    with it to fp32 rounding."""
    import numpy as np
    import unet_oracle
    from util import spectral_case
    x, mask, n_fft, hop = spectral_case(name)
    want = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_spectral.npz"))[name])
    got = unet_oracle.spectral_mask(x, mask, n_fft, hop, n_fft)
    assert torch.equal(got, want)
    if name != "ragged":                      # the per-sample python loop: keep the CPU suite short
        assert rel_l2(_stft_kernel_definition(x, mask, n_fft, hop), want) < 1e-6


# ---- the algebra behind the fused thin-end blocks (csrc/out_block.cu, csrc/conv_init.cu), checked on the oracle's resnet_block in fp64 ----
def _block_sd(p, dim, dim_out, N, seed):
    import math
    sd = {}
    if dim != N: sd[p + ".proj_in.weight"] = seeded((N, dim, 1, 1), seed + 1, 1 / math.sqrt(dim)).double()
    if dim != dim_out: sd[p + ".res_conv.weight"] = seeded((dim_out, dim, 1, 1), seed + 2, 1 / math.sqrt(dim)).double()
    if N != dim_out: sd[p + ".proj_out.weight"] = seeded((dim_out, N, 1, 1), seed + 3, 1 / math.sqrt(N)).double()
    sd[p + ".H.0.weight"] = seeded((N, N, 1, 1), seed + 4, 1 / math.sqrt(N)).double()
    sd[p + ".norm.0.gamma"] = (1 + 0.2 * seeded((1, N, 1, 1), seed + 5)).double()
    for nm, s in (("affine", 6), ("gate", 7)):
        sd[f"{p}.{nm}.0.weight"] = seeded((N, 256), seed + s, 0.05).double()
        sd[f"{p}.{nm}.0.bias"] = seeded((N,), seed + s + 10, 0.5).double()
    return sd


def test_out_block_collapses_into_two_2xN_matrices():
    """out_block.cu: with only two channels leaving the block, proj_out((x + gate * H a) / sqrt 2) + res_conv(x) is
    W1 x + W2_b a with W1 = R + P / sqrt 2 and W2_b = P diag(gate_b) H / sqrt 2 (unet.py:470-493)."""
    from oracle import unet_oracle as uo
    N, B, Fd, T = 32, 3, 5, 24
    sd = _block_sd("blk", N, 2, N, 300)
    x, emb = seeded((B, N, Fd, T), 310).double(), seeded((B, 256), 311).double()
    x[1] *= 3.0
    ref = uo.resnet_block(sd, "blk", x, emb, dim=N, dim_out=2, num_dils=1, k1x1=True, after=True)
    a = torch.nn.functional.gelu(uo.group_norm(x, sd["blk.norm.0.gamma"]) * (uo.linear(sd, "blk.affine.0", emb)[:, :, None, None] + 1))
    gate = uo.linear(sd, "blk.gate.0", emb)                                                # [B, N]
    P, R, H = sd["blk.proj_out.weight"][:, :, 0, 0], sd["blk.res_conv.weight"][:, :, 0, 0], sd["blk.H.0.weight"][:, :, 0, 0]
    A = 1 / uo.SQRT2
    W1 = A * (R + A * P)
    W2 = A * A * torch.einsum("kn,bn,nm->bkm", P, gate, H)                                 # per clip
    out = torch.einsum("kn,bnft->bkft", W1, x) + torch.einsum("bkm,bmft->bkft", W2, a)
    assert rel_l2(out, ref) < 1e-13


def test_init_block_statistics_follow_from_the_input_moments():
    """conv_init.cu: y = proj_in(x2) is a linear map of two channels, so the unbiased group variance of y (unet.py:147-163) follows from the
    five second moments of x2, and the block output is a function of the pixel's two input values plus the N x N layer."""
    from oracle import unet_oracle as uo
    N, B, Fd, T = 32, 2, 6, 20
    sd = _block_sd("blk", 2, N, N, 400)
    x2, emb = seeded((B, 2, Fd, T), 410).double(), seeded((B, 256), 411).double()
    x2[:, 1] = 0.4 * x2[:, 1] + 0.3 * x2[:, 0]
    ref = uo.resnet_block(sd, "blk", x2, emb, dim=2, dim_out=N, num_dils=1, k1x1=True)
    Wi, Wr, H = sd["blk.proj_in.weight"][:, :, 0, 0], sd["blk.res_conv.weight"][:, :, 0, 0], sd["blk.H.0.weight"][:, :, 0, 0]
    M1 = x2.sum(dim=(2, 3))                                                                # [B, 2]
    M2 = torch.einsum("bift,bjft->bij", x2, x2)                                            # [B, 2, 2]
    gcn, npg = N // 8, (N // 8) * Fd * T
    s1 = torch.einsum("nc,bc->bn", Wi, M1).reshape(B, 8, gcn).sum(-1)
    s2 = torch.einsum("ni,bij,nj->bn", Wi, M2, Wi).reshape(B, 8, gcn).sum(-1)
    std = ((s2 - s1 * s1 / npg) / (npg - 1)).sqrt()                                        # [B, 8]
    y = torch.einsum("nc,bcft->bnft", Wi, x2)
    assert torch.allclose(std, y.reshape(B, 8, -1).std(-1), rtol=1e-12)
    scale = (sd["blk.norm.0.gamma"] * (uo.linear(sd, "blk.affine.0", emb)[:, :, None, None] + 1)) / (std.repeat_interleave(gcn, 1)[:, :, None, None] + 1e-7)
    a = torch.nn.functional.gelu(y * scale)
    gate = uo.linear(sd, "blk.gate.0", emb)[:, :, None, None]
    A = 1 / uo.SQRT2
    c = Wi + uo.SQRT2 * Wr                                                                 # the epilogue's coefficient tables
    out = 0.5 * gate * torch.einsum("nm,bmft->bnft", H, a) + 0.5 * torch.einsum("nc,bcft->bnft", c, x2)
    assert rel_l2(out, ref) < 1e-13
