"""Device-resident sampler step: Philox noise generated on the GPU, device-scalar forward / step kernels, and the
step loop replayed from captured CUDA graphs -- against the numpy restatement of the generator (oracle/philox_oracle.py) and the
oracle sampler (oracle/unet_oracle.py) fed with the same noise."""
import ctypes as C

import numpy as np
import pytest
import torch

from util import rel_l2, seeded, make_oracle
from test_host import _tester_args

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [4096, 4099, 262144])
def test_philox_normal_matches_oracle(aid, cuda, L):
    import philox_oracle as po
    from aid_b200 import _lib
    lib = _lib.lib()
    seed, stream_id, clip0, draw, B = 0x1234_5678_9ABC_DEF0, 3, 5, 7, 3
    x = torch.full((B, L), 9.0, device=cuda)
    _lib.check(lib.aid_philox_normal(_lib.ptr(x), B, L, seed, stream_id, clip0, draw, 0.5, 0, None, None))
    want = 0.5 * po.batch_normals(seed, stream_id, clip0, B, draw, L)
    got = x.cpu().numpy()
    assert np.abs(got - want).max() < 2e-6 * max(1.0, np.abs(want).max())       # float32 log / sincos rounding only
    # accumulate, and the key from device scalars (what a replayed graph does): [scale, draw, stream_id, clip0] bit patterns
    sc = torch.from_numpy(np.array([np.float32(0.25).view(np.uint32), draw + 1, stream_id, clip0 + 1], dtype=np.uint32).view(np.float32)).to(cuda)
    _lib.check(lib.aid_philox_normal(_lib.ptr(x), B, L, seed, 99, 99, 99, 123.0, 1, _lib.ptr(sc), None))
    want2 = want + 0.25 * po.batch_normals(seed, stream_id, clip0 + 1, B, draw + 1, L)
    assert np.abs(x.cpu().numpy() - want2).max() < 4e-6 * max(1.0, np.abs(want2).max())
    # a clip's noise does not depend on the batch it is drawn in (rank-count independence of the sharded sampler)
    solo = torch.empty(1, L, device=cuda)
    _lib.check(lib.aid_philox_normal(_lib.ptr(solo), 1, L, seed, stream_id, clip0 + 2, draw, 0.5, 0, None, None))
    assert np.array_equal(solo.cpu().numpy()[0], got[2])
    if L >= 262144:
        n = got / 0.5
        assert abs(n.mean()) < 5e-3 and abs(n.std() - 1) < 5e-3 and abs((n ** 4).mean() - 3) < 0.05


def test_forward_and_step_with_device_scalars(aid, cuda):
    """aid_unet_forward_ds / aid_edm_step_ds read the preconditioning and (sigma, h) from device memory: same results as the
    by-value entry points."""
    from aid_b200 import _lib
    lib = _lib.lib()
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1234))
    x = seeded((2, cfg.audio_len), 5, 0.4).to(cuda)
    cn = torch.tensor([-0.7], device=cuda)
    want = net.denoise_fused(x, cn, 1.7, 0.3, 0.6)
    sc = torch.tensor([1.7, 0.3, 0.6, -0.7], device=cuda)
    out = torch.empty_like(x)
    ws = net._workspace(2, cuda)
    _lib.check(lib.aid_unet_forward_ds(net._handle, _lib.ptr(x), _lib.ptr(sc[3:]), 1, _lib.ptr(out), 2, _lib.ptr(sc), _lib.ptr(ws), ws.numel(), None), net._handle)
    assert rel_l2(out, want) < 1e-5
    n = x.numel()
    xh, d0, xb = (seeded((2, cfg.audio_len), s).to(cuda) for s in (1, 2, 3))
    a, b = torch.empty_like(x), torch.empty_like(x)
    sh = torch.tensor([0.37, -0.11], device=cuda)
    _lib.check(lib.aid_edm_step(_lib.ptr(x), _lib.ptr(xh), None, None, 0, n, 0.37, -0.11, 1, _lib.ptr(d0), _lib.ptr(xb), None, _lib.ptr(a), None))
    _lib.check(lib.aid_edm_step_ds(_lib.ptr(x), _lib.ptr(xh), None, None, 0, n, _lib.ptr(sh), 1, _lib.ptr(d0), _lib.ptr(xb), None, _lib.ptr(b), None))
    assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["inpaint", "uncond", "spectral"])
@pytest.mark.parametrize("mode", [0, 2])
def test_device_noise_sampler_graph_vs_eager_vs_oracle(aid, cuda, kind, mode):
    """Six Heun steps (11 denoiser calls) with device noise: the CUDA-graph replay equals the eager launch sequence, and both follow
    the oracle sampler driven by the numpy restatement of the same Philox stream."""
    import philox_oracle as po
    import unet_oracle
    from util import spectral_mask_rect
    cfg = aid.small_test(16384, conv_mode=mode)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    oracle = make_oracle(cfg, sd)
    args = _tester_args(aid, T=6)
    L, B = cfg.audio_len, 2
    y = seeded((B, L), 7, 0.063)
    mask = torch.ones(1, L)
    mask[..., L // 2 - 750: L // 2 + 750] = 0
    dn = aid.DeviceNoise(seed=77, stream_id=2, clip0=10)

    def run(graph):
        s = aid.Sampler(net, aid.EDM(args), args)
        s.device_noise, s.use_cuda_graph = dn, graph
        outs = []
        for _ in range(2 if graph else 1):          # the second call replays the cached graphs
            if kind == "inpaint":
                outs.append(s.predict_inpainting((y * mask).to(cuda), mask.to(cuda)))
            elif kind == "uncond":
                outs.append(s.predict_unconditional((B, L), cuda))
            else:
                s.mask = spectral_mask_rect(L, gap_ms=100).to(cuda)
                ym = s.apply_spectral_mask(y.to(cuda))
                outs.append(s.predict_spectrogram_inpainting(ym, s.mask))
        return outs

    eager = run(False)[0]
    g1, g2 = run(True)
    tol_rr = 1e-5 if mode == 0 else 2e-4            # statistic atomics move a few fp16 operand roundings in conv_mode 2
    assert rel_l2(g1, eager) < tol_rr and rel_l2(g2, eager) < tol_rr
    noise = po.PhiloxNoise(dn.seed, dn.stream_id, dn.clip0, dn.clip0 + B, L)
    edm = unet_oracle.EDMOracle()
    if kind == "inpaint":
        want = unet_oracle.sample_oracle(oracle, edm, (B, L), noise, nb_steps=6, y=y * mask, mask_s=unet_oracle.smooth_mask(mask.expand(B, -1), 50))
    elif kind == "uncond":
        want = unet_oracle.sample_oracle(oracle, edm, (B, L), noise, nb_steps=6, hpf=oracle.CQTransform.apply_hpf_DC)
    else:
        sm = spectral_mask_rect(L, gap_ms=100)
        ym = unet_oracle.spectral_mask(y, sm)
        want = unet_oracle.sample_oracle(oracle, edm, (B, L), noise, nb_steps=6, y=ym, project=unet_oracle.spectral_projection(ym, sm))
    err = rel_l2(g1, want)
    print(f"device-noise sampler ({kind}, conv_mode {mode}) vs oracle sampler on the same Philox stream: {err:.3e}")
    assert err < (1e-3 if mode == 0 else 3e-3)


def test_sharded_sampler_device_noise_is_rank_count_independent(aid, cuda):
    """One process, two 'ranks' emulated by clip offsets: clips [0,2) and [2,3) sampled separately equal the 3-clip batch."""
    from aid_b200.dist import ShardedSampler
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1234))
    args = _tester_args(aid, T=3)
    L = cfg.audio_len
    y = seeded((3, L), 7, 0.063).to(cuda)
    mask = torch.ones(1, L, device=cuda)
    mask[..., 8000:8600] = 0
    sh = ShardedSampler(aid.Sampler(net, aid.EDM(args), args), seed=5)
    full = sh.predict_inpainting(y * mask, mask)
    again = sh.predict_inpainting(y * mask, mask)
    assert rel_l2(again, full) > 1e-3                       # second call: fresh noise (call number is part of the key)
    parts = []
    for lo, hi in ((0, 2), (2, 3)):
        s = aid.Sampler(net, aid.EDM(args), args)
        s.device_noise = aid.DeviceNoise(5, stream_id=0, clip0=lo)
        parts.append(s.predict_inpainting((y * mask)[lo:hi], mask))
    assert rel_l2(torch.cat(parts), full) < 1e-5
