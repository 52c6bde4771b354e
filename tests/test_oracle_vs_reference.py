"""Pins the oracle to the reference's own code, live (build container only: needs /root/reference)."""
import sys

import pytest
import torch

from util import rel_l2, seeded, make_oracle

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, "/root/reference")
    import cqt_oracle
    cqt_oracle.install_as_cqt_nsgt_pytorch()
    from networks.unet_cqt_oct_with_projattention_adaLN_2 import Unet_CQT_oct_with_attention
    from diff_params.edm import EDM
    from testing.edm_sampler_inpainting import Sampler
    return Unet_CQT_oct_with_attention, EDM, Sampler


def _args(aid, cfg, T):
    sys.path.insert(0, __file__.rsplit("/", 1)[0] + "/golden")
    from make_golden import TESTER
    a = cfg.to_args()
    a.update(aid.AttrDict.wrap(TESTER))
    a["tester"]["T"] = T
    return a


def test_schema_and_forward(aid, ref):
    RefNet, _, _ = ref
    cfg = aid.NetConfig(audio_len=32768, Ns=[8, 16, 16, 24, 24, 32, 40], num_dils=[2, 1, 3, 2, 2, 3, 2],
                        attention_layers=[0, 0, 0, 1, 0, 1, 1, 1])
    net = RefNet(cfg.to_args(), "cpu")
    rsd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in rsd.items()} == dict(aid.schema_from_lib(cfg))
    sd = aid.random_state_dict(cfg, seed=99)
    net.load_state_dict(sd, strict=True)
    x = seeded((2, cfg.audio_len), 4, 0.7)
    cn = torch.tensor([[-1.1], [0.02]])
    with torch.no_grad():
        want = net(x, cn)
    assert rel_l2(make_oracle(cfg, sd)(x, cn), want) < 1e-6


def test_sampler_mirror_equals_reference_sampler(aid, ref):
    """Our Sampler (host logic, torch ops on CPU) against the reference Sampler with the same cheap fake denoiser."""
    _, RefEDM, RefSampler = ref
    cfg = aid.small_test(16384)

    class FakeNet(torch.nn.Module):
        class _C:
            @staticmethod
            def apply_hpf_DC(x):
                return x - x.mean(-1, keepdim=True)
        CQTransform = _C()

        def forward(self, x, cnoise):
            return torch.tanh(3.0 * x) * (1.0 + 0.1 * cnoise)

    net = FakeNet()
    for T, order in ((35, 2), (5, 1)):
        args = _args(aid, cfg, T)
        args["tester"]["order"] = order
        y = seeded((3, 4096), 1, 0.063)
        mask = torch.ones(1, 4096)
        mask[..., 1000:1400] = 0
        mask[..., 3000:3100] = 0
        torch.manual_seed(5)
        want_i = RefSampler(net, RefEDM(args), args).predict_inpainting(y * mask, mask)
        torch.manual_seed(5)
        got_i = aid.Sampler(net, aid.EDM(args), args).predict_inpainting(y * mask, mask)
        assert torch.equal(got_i, want_i)
        torch.manual_seed(6)
        want_u = RefSampler(net, RefEDM(args), args).predict_unconditional((2, 4096), "cpu")
        torch.manual_seed(6)
        got_u = aid.Sampler(net, aid.EDM(args), args).predict_unconditional((2, 4096), "cpu")
        assert torch.equal(got_u, want_u)
    s = aid.Sampler(net, aid.EDM(args), args)
    assert torch.equal(s.prepare_smooth_mask(mask.expand(3, -1), 50), RefSampler(net, RefEDM(args), args).prepare_smooth_mask(mask.expand(3, -1), 50))


def test_reference_checkpoint_loader_works_on_our_module(aid, ref):
    """tester_inpainting.py:195-202 -> utils/training_utils.py:214-382 load_state_dict(state_dict, ema=network): the
    reference's own multi-strategy loader drives our nn.Module unchanged (strict, then shape-matched partial load)."""
    import utils.training_utils as t_utils
    cfg = aid.small_test(16384)
    sd = aid.random_state_dict(cfg, seed=21)
    net = aid.Unet_CQT_oct_with_attention(cfg.to_args(), "cpu")
    assert t_utils.load_state_dict({"it": 5, "ema": sd}, ema=net, log=False) is True
    assert torch.equal(net.state_dict()["ups.3.1.H.2.weight"], sd["ups.3.1.H.2.weight"])
    # attempt 3: a checkpoint with one wrongly shaped tensor still loads everything else
    bad = dict(sd)
    bad["downs.0.1.weight"] = torch.zeros(3, 3)
    net2 = aid.Unet_CQT_oct_with_attention(cfg.to_args(), "cpu")
    assert t_utils.load_state_dict({"ema": bad, "network": bad}, network=net2, ema=net2, log=False) is True
    assert torch.equal(net2.state_dict()["middle.0.1.qk.weight" if False else "middle.0.1.attn_block.qk.weight"], sd["middle.0.1.attn_block.qk.weight"])


def test_spectrogram_inpainting_mirror_equals_reference_sampler(aid, ref):
    """predict_spectrogram_inpainting / apply_spectral_mask (sampler.py:271-290, 348-364), CPU tensors, same fake denoiser."""
    from test_host import _FakeNet
    from util import spectral_mask_rect
    _, RefEDM, RefSampler = ref
    cfg = aid.small_test(16384)
    args = _args(aid, cfg, 6)
    net = _FakeNet()
    L = 8192
    y = seeded((2, L), 2, 0.063)
    m = spectral_mask_rect(L, gap_ms=100)
    rs, s = RefSampler(net, RefEDM(args), args), aid.Sampler(net, aid.EDM(args), args)
    rs.mask = s.mask = m
    ym = rs.apply_spectral_mask(y)
    assert torch.equal(s.apply_spectral_mask(y), ym)
    torch.manual_seed(8)
    want = rs.predict_spectrogram_inpainting(ym, m)
    torch.manual_seed(8)
    got = s.predict_spectrogram_inpainting(ym, m)
    assert torch.equal(got, want)


@pytest.mark.parametrize("consistency", [True, False])
def test_guidance_branch_mirror_equals_reference_sampler(aid, ref, consistency):
    """xi = 0.25 (sampler.py:55-113), batch 1 -- the only batch size the reference's autograd.grad call accepts -- with and
    without the projection, plus the rid=True diagnostics."""
    from test_host import _FakeNet
    _, RefEDM, RefSampler = ref
    cfg = aid.small_test(16384)
    args = _args(aid, cfg, 6)
    args["tester"]["posterior_sampling"]["xi"] = 0.25
    args["tester"]["data_consistency"]["use"] = consistency
    net = _FakeNet()
    L = 4096
    y = seeded((1, L), 1, 0.063)
    mask = torch.ones(1, L)
    mask[..., 1000:1400] = 0
    for rid in (False, True):
        torch.manual_seed(5)
        want = RefSampler(net, RefEDM(args), args, rid).predict_inpainting(y * mask, mask)
        torch.manual_seed(5)
        got = aid.Sampler(net, aid.EDM(args), args, rid).predict_inpainting(y * mask, mask)
        if rid:
            assert len(got) == len(want) == 8
            for g, w in zip(got, want):
                assert rel_l2(g, w) < 1e-6
        else:
            assert rel_l2(got, want) < 1e-6


def test_mask_builders_equal_the_reference_tester(aid, ref):
    """prepare_mask / prepare_spectral_mask against testing.tester_inpainting.Tester's methods, called unbound on a namespace
    that carries what they read (args, device)."""
    import importlib
    import types
    from unittest import mock
    Tester = None
    for _ in range(30):     # the tester imports plotting / logging packages that are not installed here: stub what is missing
        try:
            Tester = importlib.import_module("testing.tester_inpainting").Tester
            break
        except ModuleNotFoundError as e:
            sys.modules[e.name] = mock.MagicMock()
    assert Tester is not None
    for L in (65536, 184184):
        a = _args(aid, aid.small_test(16384), 6)
        a["exp"]["audio_len"] = L
        a["tester"]["inpainting"] = aid.AttrDict.wrap({"mask_mode": "long", "long": {"gap_length": 1500, "start_gap_idx": "None"},
                                                       "short": {"num_gaps": 4, "gap_length": 25, "start_gap_idx": "None"}})
        ns = types.SimpleNamespace(args=a, device="cpu")
        assert torch.equal(aid.prepare_mask(a), Tester.prepare_mask(ns))
        a["tester"]["inpainting"]["long"]["start_gap_idx"] = 250
        assert torch.equal(aid.prepare_mask(a), Tester.prepare_mask(ns))
        a["tester"]["inpainting"]["mask_mode"] = "short"
        torch.manual_seed(1)
        want = Tester.prepare_mask(ns)
        torch.manual_seed(1)
        assert torch.equal(aid.prepare_mask(a), want)
        assert torch.equal(aid.prepare_spectral_mask(a), Tester.prepare_spectral_mask(ns))
        a["tester"]["spectrogram_inpainting"]["time_start_idx"] = 1000
        assert torch.equal(aid.prepare_spectral_mask(a), Tester.prepare_spectral_mask(ns))


def test_eight_octave_44k_network_schema_and_forward(aid, ref):
    """conf/network/paper_1912_unet_cqt_oct_attention_44k_2.yaml (8 octaves, fs = 44100, attention from level 5, dilations up to 2^7):
    the full-width schema (795 tensors, 242.25 M parameters) equals the reference module's, and the oracle equals the reference's
    forward on a narrowed copy of the same topology at the reference's trained length 184184 (conf/exp/musicnet44k_4s.yaml)."""
    RefNet, _, _ = ref
    full = aid.paper_44k()
    sch = aid.schema_from_lib(full)
    assert len(sch) == 795 and abs(sum(torch.Size(s).numel() for _, s in sch) / 1e6 - 242.25) < 0.01
    cfg = aid.paper_44k(184184)
    cfg.Ns, cfg.num_dils = [8, 8, 16, 16, 16, 24, 24, 32], [2, 3, 4, 5, 6, 7, 8, 8]
    net = RefNet(cfg.to_args(), "cpu")
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == dict(aid.schema_from_lib(cfg))
    sd = aid.random_state_dict(cfg, seed=44)
    net.load_state_dict(sd, strict=True)
    x = seeded((1, cfg.audio_len), 4, 0.7)
    cn = torch.tensor([[-0.4]])
    with torch.no_grad():
        want = net(x, cn)
    assert rel_l2(make_oracle(cfg, sd)(x, cn), want) < 1e-6


def test_predict_resample_mirror_equals_reference_sampler(aid, ref):
    """predict_resample(y, shape, degradation) (sampler.py:164-173): the generic conditional entry point -- the caller supplies the
    degradation and, for xi = 0, must have set proj_convex_set (get_score reads it, sampler.py:145); without one both raise."""
    from test_host import _FakeNet
    _, RefEDM, RefSampler = ref
    cfg = aid.small_test(16384)
    args = _args(aid, cfg, 6)
    net = _FakeNet()
    L = 4096
    y = seeded((2, L), 3, 0.063)
    mask = torch.ones(1, L)
    mask[..., 700:1500] = 0
    deg = lambda x: mask * x
    rs, s = RefSampler(net, RefEDM(args), args), aid.Sampler(net, aid.EDM(args), args)
    with pytest.raises(AttributeError):
        rs.predict_resample(y * mask, (2, L), deg)
    with pytest.raises(AttributeError):
        s.predict_resample(y * mask, (2, L), deg)
    for smp in (rs, s):
        smp.proj_convex_set = lambda x: mask * (y * mask) + (1 - mask) * x
    torch.manual_seed(9)
    want = rs.predict_resample(y * mask, (2, L), deg)
    torch.manual_seed(9)
    got = s.predict_resample(y * mask, (2, L), deg)
    assert torch.equal(got, want)
    # xi > 0 at batch 1: guidance through the caller's degradation, no projection needed when data consistency is off
    args["tester"]["posterior_sampling"]["xi"] = 0.25
    args["tester"]["data_consistency"]["use"] = False
    rs, s = RefSampler(net, RefEDM(args), args), aid.Sampler(net, aid.EDM(args), args)
    torch.manual_seed(10)
    want = rs.predict_resample(y[:1] * mask, (1, L), deg)
    torch.manual_seed(10)
    got = s.predict_resample(y[:1] * mask, (1, L), deg)
    assert rel_l2(got, want) < 1e-6
