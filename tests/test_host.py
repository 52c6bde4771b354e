"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, schema, config parsing,
CQT band plan (C++) against the oracle's, sampler host logic, error behaviour.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from util import rel_l2, seeded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol(aid):
    from aid_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "aid_b200.h")).read()
    declared = set(re.findall(r"\b(aid_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/aid_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_config_from_reference_style_args(aid):
    cfg = aid.paper_22k(262144)
    cfg2 = aid.NetConfig.from_args(cfg.to_args())
    assert cfg2.as_dict() == cfg.as_dict()
    bad = cfg.to_args()
    bad["network"]["use_fencoding"] = True
    with pytest.raises(NotImplementedError):
        aid.NetConfig.from_args(bad)


def test_unsupported_length_is_a_clear_error(aid):
    with pytest.raises(Exception, match="audio_len must be even"):
        aid.schema_from_lib(aid.NetConfig(audio_len=184185))
    # the reference's trained length (conf/exp/maestro22k_8s.yaml:52) is supported (Bluestein FFT), T0 = 2048 frames
    assert len(aid.schema_from_lib(aid.NetConfig(audio_len=184184))) == 662


def test_module_surface_and_state_dict_roundtrip(aid):
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg.to_args(), "cpu")
    sd = aid.random_state_dict(cfg, seed=3)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd, strict=True)
    assert torch.equal(net.state_dict()["downs.2.2.H.1.weight"], sd["downs.2.2.H.1.weight"])
    assert hasattr(net.CQTransform, "apply_hpf_DC") and hasattr(net.CQTransform, "fwd") and hasattr(net.CQTransform, "bwd")
    with pytest.raises(RuntimeError):
        net.load_state_dict({k: v for k, v in sd.items() if "gate2" not in k}, strict=True)
    # default init keeps the reference's property: gates ~1e-7, gammas 1 (unet.py:599-600, 141)
    fresh = aid.Unet_CQT_oct_with_attention(cfg, "cpu").state_dict()
    assert fresh["downs.0.2.gate.0.weight"].abs().max() < 1e-6 and torch.all(fresh["downs.0.2.norm.0.gamma"] == 1)


def test_no_cpu_fallback(aid):
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg, "cpu")
    from aid_b200 import _lib
    with pytest.raises(_lib.AidError, match="CUDA device only"):
        net(torch.zeros(1, cfg.audio_len), torch.zeros(1, 1))


def test_workspace_plan_is_host_only_and_grows_with_batch(aid):
    from aid_b200 import _lib
    L = _lib.lib()
    cfg = aid.paper_22k(262144)
    h = C.c_void_p()
    c = cfg.to_c()
    _lib.check(L.aid_create(C.byref(c), 0, C.byref(h)))
    try:
        n1, n2, n32 = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.check(L.aid_workspace_bytes(h, 1, C.byref(n1)), h)
        _lib.check(L.aid_workspace_bytes(h, 2, C.byref(n2)), h)
        _lib.check(L.aid_workspace_bytes(h, 32, C.byref(n32)), h)
        assert n1.value < n2.value < n32.value < 100e9  # fits one B200 (180 GB) with room for the weights
        assert L.aid_num_weights(h) == 662
    finally:
        L.aid_destroy(h)


@pytest.mark.parametrize("L", [16384, 65536, 184184, 262144])
def test_cqt_band_plan_matches_oracle(aid, L):
    """The C++ band plan (csrc/cqt_plan.hpp) and the oracle's numpy plan are two writings of one definition."""
    import cqt_oracle
    from aid_b200 import _lib
    Lb = _lib.lib()
    cfg = aid.small_test(L)
    p = cqt_oracle.CQTPlan(cfg.num_octs, cfg.bins_per_oct, ("kaiser", cfg.beta), cfg.sample_rate, L)
    h = C.c_void_p()
    c = cfg.to_c()
    _lib.check(Lb.aid_create(C.byref(c), 0, C.byref(h)))
    try:
        K, nw = C.c_int32(), C.c_int32()
        _lib.check(Lb.aid_cqt_plan(h, C.byref(K), C.byref(nw), None, None, None, None, None, None))
        cen, lg, wo = np.zeros(K.value, np.int32), np.zeros(K.value, np.int32), np.zeros(K.value, np.int32)
        win, dual, hh = np.zeros(nw.value, np.float32), np.zeros(nw.value, np.float32), np.zeros(L, np.float32)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(Lb.aid_cqt_plan(h, None, None, P(cen), P(lg), P(wo), P(win), P(dual), P(hh)))
        offs, frames = (C.c_int64 * 8)(), (C.c_int32 * 7)()
        _lib.check(Lb.aid_cqt_layout(h, 1, offs, frames))
    finally:
        Lb.aid_destroy(h)
    assert K.value == p.K and list(frames) == p.size_per_oct
    assert np.array_equal(cen, p.centre[1:p.K + 1]) and np.array_equal(lg, p.Lg[1:p.K + 1])
    assert np.allclose(hh, p.Hhpf, atol=1e-6)
    for k in (0, 1, p.K // 2, p.K - 2, p.K - 1):
        j = k + 1
        o = k // cfg.bins_per_oct
        assert np.allclose(win[wo[k]:wo[k] + lg[k]], p.g[j], atol=1e-6)
        assert np.allclose(dual[wo[k]:wo[k] + lg[k]], p.gd[j] * p.size_per_oct[o], rtol=1e-5, atol=1e-9)


class _FakeNet(torch.nn.Module):
    class _C:
        @staticmethod
        def apply_hpf_DC(x):
            return x - x.mean(-1, keepdim=True)
    CQTransform = _C()

    def forward(self, x, cnoise):
        return torch.tanh(3.0 * x) * (1.0 + 0.1 * cnoise)


def _tester_args(aid, T=35, order=2):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    a = aid.AttrDict.wrap({
        "tester": {"T": T, "order": order, "filter_out_cqt_DC_Nyq": True, "posterior_sampling": {"xi": 0, "norm": 2, "smoothl1_beta": 1},
                   "data_consistency": {"use": True, "type": "always", "smooth": True, "hann_size": 50},
                   "spectrogram_inpainting": {"stft": {"window": "hann", "n_fft": 1024, "hop_length": 256, "win_length": 1024},
                                              "time_mask_length": 2000, "time_start_idx": "None", "min_masked_freq": 300,
                                              "max_masked_freq": 2000},
                   "diff_params": {"same_as_training": False, "sigma_data": 0.063, "sigma_min": 1e-4, "sigma_max": 1, "ro": 13,
                                   "Schurn": 10, "Snoise": 1.0, "Stmin": 0, "Stmax": 50}},
        "diff_params": {"sigma_data": 0.063, "sigma_min": 1e-5, "sigma_max": 10, "ro": 13, "Schurn": 5, "Snoise": 1, "Stmin": 0, "Stmax": 50},
        "exp": {"audio_len": 8192, "sample_rate": 22050},
    })
    return a


def test_sampler_host_logic_against_oracle(aid):
    """69 denoiser evaluations for T=35 and the same trajectory as oracle/sample_oracle."""
    import unet_oracle
    args = _tester_args(aid)
    net = _FakeNet()
    calls = []
    net.register_forward_hook(lambda m, i, o: calls.append(1))
    s = aid.Sampler(net, aid.EDM(args), args)
    assert s.diff_params.sigma_max == 1 and s.diff_params.Schurn == 10  # tester overrides applied (sampler.py:43-53)
    y = seeded((2, 8192), 1, 0.063)
    mask = torch.ones(1, 8192)
    mask[..., 4000:4600] = 0

    def stream(shape):
        while True:
            yield torch.randn(shape)

    torch.manual_seed(11)
    got = s.predict_inpainting(y * mask, mask)
    assert len(calls) == 69
    torch.manual_seed(11)
    want = unet_oracle.sample_oracle(net, unet_oracle.EDMOracle(), (2, 8192), stream((2, 8192)), nb_steps=35, y=y * mask,
                                     mask_s=unet_oracle.smooth_mask(mask.expand(2, -1), 50))
    assert rel_l2(got, want) < 1e-6
    # known samples are restored by the replacement method (the last Euler step lands on the projected estimate)
    assert torch.allclose(got[:, :3900], (y * mask)[:, :3900], atol=1e-6)
    torch.manual_seed(12)
    got_u = s.predict_unconditional((2, 8192), "cpu")
    torch.manual_seed(12)
    want_u = unet_oracle.sample_oracle(net, unet_oracle.EDMOracle(), (2, 8192), stream((2, 8192)), nb_steps=35, hpf=net.CQTransform.apply_hpf_DC)
    assert rel_l2(got_u, want_u) < 1e-6


def test_sampler_guidance_is_refused_on_the_forward_only_denoiser(aid):
    """xi > 0 needs the denoiser's VJP: a model exposing the CUDA path's `denoise_fused` is refused before any work."""
    class ForwardOnly(_FakeNet):
        def denoise_fused(self, *a, **k):
            raise AssertionError("must not be reached")

    args = _tester_args(aid)
    args["tester"]["posterior_sampling"]["xi"] = 0.25
    s = aid.Sampler(ForwardOnly(), aid.EDM(args), args)
    with pytest.raises(NotImplementedError, match="xi"):
        s.predict_inpainting(torch.zeros(1, 4096), torch.ones(1, 4096))


@pytest.mark.parametrize("consistency", [True, False])
def test_sampler_guidance_host_logic(aid, consistency):
    """sampler.py:55-113 with a differentiable stand-in denoiser: equals the oracle loop at batch 1 (the reference's only
    working batch size); at batch 2 every clip gets what it gets alone (per-clip norm and step size)."""
    import unet_oracle
    L = 8192
    args = _tester_args(aid, T=6)
    args["tester"]["posterior_sampling"]["xi"] = 0.25
    args["tester"]["data_consistency"]["use"] = consistency
    net = _FakeNet()
    y = seeded((2, L), 1, 0.063)
    mask = torch.ones(1, L)
    mask[..., 4000:4600] = 0
    outs = []
    for b in range(2):
        yb = y[b:b + 1] * mask

        def stream():
            g = torch.Generator().manual_seed(50 + b)
            while True:
                yield torch.randn(1, L, generator=g)

        s = aid.Sampler(net, aid.EDM(args), args)
        s.noise_source = stream()
        got = s.predict_inpainting(yb, mask)
        want = unet_oracle.sample_oracle(net, unet_oracle.EDMOracle(), (1, L), stream(), nb_steps=6, y=yb,
                                         mask_s=unet_oracle.smooth_mask(mask, 50) if consistency else None,
                                         hpf=net.CQTransform.apply_hpf_DC,       # filter_out_cqt_DC_Nyq acts on this branch (sampler.py:62)
                                         guidance={"xi": 0.25, "degradation": lambda x: mask * x, "audio_len": 8192})
        assert rel_l2(got, want) < 1e-5
        outs.append(got)

    def both():
        gens = [torch.Generator().manual_seed(50 + b) for b in range(2)]
        while True:
            yield torch.cat([torch.randn(1, L, generator=g) for g in gens])

    s = aid.Sampler(net, aid.EDM(args), args)
    s.noise_source = both()
    got2 = s.predict_inpainting(y * mask, mask)
    assert rel_l2(got2, torch.cat(outs)) < 1e-5
    s.rid, s.noise_source = True, both()
    rid = s.predict_inpainting(y * mask, mask)
    assert len(rid) == 8 and torch.equal(rid[0], got2) and rid[1].shape == (6, 2, L) and rid[7].shape == (7,)


def test_smooth_mask_matches_the_reference_loop(aid):
    """sampler.py:302-325 restated as the literal loop."""
    args = _tester_args(aid)
    s = aid.Sampler(_FakeNet(), aid.EDM(args), args)
    mask = torch.ones(2, 3000)
    mask[:, 500:900] = 0
    mask[:, 2000:2010] = 0
    size = 50
    hann = torch.hann_window(size * 2)
    m, new, prev = mask[0], mask[0].clone(), 1
    for i in range(len(m)):
        if m[i] != prev:
            if m[i] == 0:
                new[i - size:i] = hann[size:]
            if m[i] == 1:
                new[i:i + size] = hann[:size]
        prev = m[i]
    assert torch.equal(s.prepare_smooth_mask(mask, size), new.unsqueeze(0).expand(2, -1))


def test_spectrogram_inpainting_host_logic(aid):
    """sampler.py:271-290, 348-364 on CPU tensors: the degradation equals the reference's golden output bit for bit, and the
    sampling loop with the projection y + x - S(x) equals the oracle loop."""
    import numpy as np
    import unet_oracle
    from util import spectral_case, spectral_mask_rect
    args = _tester_args(aid, T=5)
    s = aid.Sampler(_FakeNet(), aid.EDM(args), args)
    x, mask, n_fft, hop = spectral_case("ragged")
    s.mask = mask
    want = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "golden_spectral.npz"))["ragged"])
    assert torch.equal(s.apply_spectral_mask(x), want)
    L = 8192
    y = seeded((2, L), 2, 0.063)
    m = spectral_mask_rect(L, gap_ms=100)
    s.mask = m
    y_masked = s.apply_spectral_mask(y)

    def stream(shape):
        while True:
            yield torch.randn(shape)

    torch.manual_seed(21)
    got = s.predict_spectrogram_inpainting(y_masked, m)
    torch.manual_seed(21)
    want = unet_oracle.sample_oracle(_FakeNet(), unet_oracle.EDMOracle(), (2, L), stream((2, L)), nb_steps=5, y=y_masked,
                                     project=unet_oracle.spectral_projection(y_masked, m))
    assert rel_l2(got, want) < 1e-6
    # (no consistency property is asserted: masking an STFT is not idempotent, S(S(x)) != S(x), so y + x - S(x) is not an
    #  exact projection -- the reference's own comment at sampler.py:361 says as much)
    # another mode afterwards does not inherit the spectral projection
    tm = torch.ones(1, L); tm[..., 4000:4300] = 0
    torch.manual_seed(22)
    a = s.predict_inpainting(y * tm, tm)
    torch.manual_seed(22)
    b = aid.Sampler(_FakeNet(), aid.EDM(args), args).predict_inpainting(y * tm, tm)
    assert torch.equal(a, b)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver runs it next to our arm): one JSON line with the contract's keys, produced by the
    CPU oracle without touching CUDA; under torchrun only rank 0 prints, the other ranks exit 0 without work."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--len", "16384", "--steps", "1", "--warmup", "0"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["config"]["audio_len"] == 16384
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "networks", "unet_cqt_oct_with_projattention_adaLN_2.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")      # oracle/_ref is placed by build() when it can be
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 1 and line["warmup"] == 0 and line["config"]["batch_per_gpu"] == 32
    # the arm is the checker's side only: it must not import the product package (and with it libaid_b200.so)
    probe = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--len', '16384', '--steps', '1', '--warmup', '0'];"
             f"runpy.run_path({os.path.join(ROOT, 'bench.py')!r}, run_name='__main__');"
             "bad = [m for m in sys.modules if m.startswith('aid_b200') or m.startswith('audio-inpainting')]; assert not bad, bad")
    rp = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, env=env, timeout=600)
    assert rp.returncode == 0, rp.stderr[-2000:]
    r1 = subprocess.run(cmd, capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=600)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_mask_builders(aid):
    """tester_inpainting.py:231-296 restated: centred / placed long gap, random short gaps, the spectrogram rectangle."""
    from util import spectral_mask_rect
    a = _tester_args(aid)
    a["exp"] = aid.AttrDict.wrap({"audio_len": 65536, "sample_rate": 22050})
    a["tester"]["inpainting"] = aid.AttrDict.wrap({"mask_mode": "long", "long": {"gap_length": 300, "start_gap_idx": "None"},
                                                   "short": {"num_gaps": 4, "gap_length": 25, "start_gap_idx": "None"}})
    m = aid.prepare_mask(a)
    assert m.shape == (1, 65536) and int((m == 0).sum()) == 6615 and m[0, 32768 - 3307] == 0 and m[0, 32768 - 3308] == 1
    a["tester"]["inpainting"]["long"]["start_gap_idx"] = 100
    assert int(torch.nonzero(aid.prepare_mask(a)[0] == 0)[0]) == 2205
    a["tester"]["inpainting"]["mask_mode"] = "short"
    torch.manual_seed(3)
    ms = aid.prepare_mask(a)
    torch.manual_seed(3)
    starts = torch.randint(0, 65536 - 551, (4,))
    want = torch.ones(1, 65536)
    for st in starts:
        want[..., st:st + 551] = 0
    assert torch.equal(ms, want)
    sm = aid.prepare_spectral_mask(a)
    assert torch.equal(sm, spectral_mask_rect(65536)) and sm.shape == (513, 1 + (65536 + 1024) // 256)
    s = aid.Sampler(_FakeNet(), aid.EDM(a), a)      # the mask fits the sampler's STFT of a clip of that length
    s.mask = sm
    assert s.apply_spectral_mask(torch.zeros(1, 65536)).shape == (1, 65536)
