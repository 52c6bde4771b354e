"""Input gradient (VJP) of the CUDA denoiser -- the reference's reconstruction-guidance branch (sampler.py:57-113,
`torch.autograd.grad(norm, x)` through edm.py:133-148 and unet.py:730-845): every backward piece against torch autograd of the
oracle's forward, the whole network against autograd through the oracle, and a guided sampling trajectory (xi = 0.25, batch 1)
against the run of the reference's own Sampler (tests/golden/make_golden_guided.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from util import rel_l2, seeded, make_oracle
from test_host import _tester_args

pytestmark = pytest.mark.gpu


def _lib():
    from aid_b200 import _lib
    return _lib, _lib.lib()


@pytest.mark.parametrize("up", [0, 1])
@pytest.mark.parametrize("T", [16, 64, 250])
def test_resample_adjoint(cuda, up, T):
    import unet_oracle
    _l, L = _lib()
    B, Cc, F = 2, 3, 5
    x = seeded((B, Cc, F, T), 1).requires_grad_()
    y = unet_oracle.up_t(x) if up else unet_oracle.down_t(x)
    g = seeded(tuple(y.shape), 2)
    want = torch.autograd.grad(y, x, g)[0]
    gx, gd = torch.empty(B, Cc, F, T, device=cuda), g.to(cuda)      # named device tensors: a temporary would be freed before the call
    _l.check(L.aid_op_resample_adj(_l.ptr(gd), B, Cc, F, T, up, _l.ptr(gx), None))
    assert rel_l2(gx, want) < 1e-6


@pytest.mark.parametrize("gelu", [0, 1])
def test_groupnorm_act_backward(cuda, gelu):
    import unet_oracle
    _l, L = _lib()
    B, Cc, F, T = 2, 24, 6, 50
    x = seeded((B, Cc, F, T), 3, 1.7).requires_grad_()
    gamma, aff = 1 + 0.3 * seeded((Cc,), 4), 0.4 * seeded((Cc,), 5)
    y = unet_oracle.group_norm(x, gamma.view(1, Cc, 1, 1)) * (aff.view(1, Cc, 1, 1) + 1)
    if gelu:
        y = Fn.gelu(y)
    g = seeded(tuple(y.shape), 6)
    want = torch.autograd.grad(y, x, g)[0]
    gx = torch.empty(B, Cc, F, T, device=cuda)
    scr = torch.empty(B * 24, dtype=torch.float64, device=cuda)
    gd, xd, gam, af = g.to(cuda), x.detach().to(cuda), gamma.to(cuda), aff.to(cuda)
    _l.check(L.aid_op_groupnorm_act_bwd(_l.ptr(gd), _l.ptr(xd), _l.ptr(gam), _l.ptr(af), B, Cc, F, T, gelu, _l.ptr(gx), _l.ptr(scr), None))
    assert rel_l2(gx, want) < 2e-5


@pytest.mark.parametrize("case", [(2, 16, 1, 1, 1), (16, 2, 1, 1, 1), (8, 16, 1, 1, 1), (16, 8, 1, 1, 1), (2, 16, 5, 3, 1), (16, 16, 5, 3, 4),
                                  (24, 40, 1, 1, 1), (32, 32, 5, 3, 16), (2, 96, 5, 3, 1)])
def test_conv_backward_input(cuda, case):
    _l, L = _lib()
    Cin, Cout, KF, KT, dil = case
    B, F, T = 2, 40, 64
    x = seeded((B, Cin, F, T), 1).double().requires_grad_()
    w = (seeded((Cout, Cin, KF, KT), 2) / (Cin * KF * KT) ** 0.5)
    y = Fn.conv2d(x, w.double(), padding="same", dilation=(dil, 1))
    g = seeded(tuple(y.shape), 3)
    want = torch.autograd.grad(y, x, g.double())[0]
    gx, gd, wd = torch.empty(B, Cin, F, T, device=cuda), g.to(cuda), w.to(cuda)
    _l.check(L.aid_op_conv2d_bwd_input(_l.ptr(gd), _l.ptr(wd), B, Cin, Cout, F, T, KF, KT, dil, _l.ptr(gx), None))
    assert rel_l2(gx, want) < 2e-6


@pytest.mark.parametrize("shape", [(2, 8, 40, 64), (1, 8, 24, 16), (1, 4, 33, 100)])
def test_attention_core_backward(cuda, shape):
    _l, L = _lib()
    B, H, F, T = shape
    h = seeded((B, H, F, T), 1).requires_grad_()
    qk = seeded((B, 2 * H * F, T), 2, 0.7).requires_grad_()
    q = qk.view(B, H, 2 * F, T)[:, :, :F].permute(0, 1, 3, 2)      # [B,H,T,F]
    k = qk.view(B, H, 2 * F, T)[:, :, F:].permute(0, 1, 3, 2)
    v = h.permute(0, 1, 3, 2)
    attn = torch.softmax(torch.einsum("bhtd,bhsd->bhts", q, k) * F ** -0.5, dim=-1)
    o = torch.einsum("bhts,bhsf->bhtf", attn, v).permute(0, 1, 3, 2)  # [B,H,F,T]
    g = seeded(tuple(o.shape), 3)
    wh, wqk = torch.autograd.grad(o, (h, qk), g)
    gh, gqk = torch.empty(B, H, F, T, device=cuda), torch.empty(B, 2 * H * F, T, device=cuda)
    scr = torch.empty(2 * B * H * T * T * 4 + 4096, dtype=torch.uint8, device=cuda)
    hd, qkd, gd = h.detach().to(cuda), qk.detach().to(cuda), g.to(cuda)
    _l.check(L.aid_op_attention_bwd(_l.ptr(hd), _l.ptr(qkd), _l.ptr(gd), B, H, F, T, _l.ptr(gh), _l.ptr(gqk), _l.ptr(scr), scr.numel(), None))
    assert rel_l2(gh, wh) < 1e-5 and rel_l2(gqk, wqk) < 1e-5


@pytest.mark.parametrize("L", [16384, 184184])
def test_cqt_adjoints(aid, cuda, L):
    """<fwd(x), G> == <x, fwd^T G> and the same for the synthesis, against torch autograd of the CQT restatement."""
    import cqt_oracle
    _l, Lb = _lib()
    cfg = aid.small_test(L)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=3))
    net._ensure_weights(cuda)
    orc = cqt_oracle.CQT_nsgt(cfg.num_octs, cfg.bins_per_oct, "oct", ("kaiser", cfg.beta), fs=cfg.sample_rate, audio_len=L)
    B = 2
    offs, frames = net.CQTransform._layout(B)
    ws = net.CQTransform._ws(B, cuda)
    # analysis adjoint
    x = seeded((B, 1, L), 1).requires_grad_()
    X = orc.fwd(x)
    G = [torch.complex(seeded(tuple(c.shape), 10 + i), seeded(tuple(c.shape), 30 + i)) for i, c in enumerate(X)]
    val = sum((torch.view_as_real(a) * torch.view_as_real(g)).sum() for a, g in zip(X, G))
    want = torch.autograd.grad(val, x)[0][:, 0]
    gcoef = torch.empty(offs[-1], device=cuda)
    for i, g in enumerate(G):
        dst = gcoef[offs[i]:offs[i + 1]].view(B, 2, cfg.bins_per_oct, frames[i])
        dst[:, 0].copy_(g[:, 0].real); dst[:, 1].copy_(g[:, 0].imag)
    gx = torch.empty(B, L, device=cuda)
    _l.check(Lb.aid_cqt_fwd_vjp(net._handle, _l.ptr(gcoef), _l.ptr(gx), B, _l.ptr(ws), ws.numel(), None), net._handle)
    assert rel_l2(gx, want) < 1e-5
    # synthesis adjoint
    coefs = [torch.complex(seeded(tuple(c.shape), 50 + i), seeded(tuple(c.shape), 70 + i)).requires_grad_() for i, c in enumerate(X)]
    y = orc.bwd(coefs)
    gy = seeded(tuple(y.shape), 5)
    wants = torch.autograd.grad(y, coefs, gy)
    gcoef2 = torch.empty(offs[-1], device=cuda)
    gyd = gy[:, 0, :L].contiguous().to(cuda)
    _l.check(Lb.aid_cqt_bwd_vjp(net._handle, _l.ptr(gyd), _l.ptr(gcoef2), B, _l.ptr(ws), ws.numel(), None), net._handle)
    for i, w in enumerate(wants):
        got = gcoef2[offs[i]:offs[i + 1]].view(B, 2, cfg.bins_per_oct, frames[i])
        # torch's gradient of a real loss with respect to a complex leaf is dL/dRe + i dL/dIm
        assert rel_l2(got[:, 0], w[:, 0].real) < 1e-5 and rel_l2(got[:, 1], w[:, 0].imag) < 1e-5, i


@pytest.mark.parametrize("mode", [0, 2])
def test_denoiser_vjp_vs_oracle_autograd(aid, cuda, mode):
    """grad_x of <g, D(x; sigma)> for the whole preconditioned denoiser (edm.py:133-148 + unet.py:730-845), batch 2, small network."""
    import unet_oracle
    cfg = aid.small_test(16384, conv_mode=mode)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    orc = make_oracle(cfg, sd)
    edm_o = unet_oracle.EDMOracle()
    edm = aid.EDM(_tester_args(aid))
    aid.Sampler(net, edm, _tester_args(aid))          # applies the tester's diff_params (sigma_data etc.)
    x = seeded((2, cfg.audio_len), 0, 0.4)
    g = seeded((2, cfg.audio_len), 9)
    sigma = torch.tensor([0.37])
    xo = x.clone().requires_grad_()
    want_out = edm_o.denoiser(xo, orc.differentiable, sigma)
    want = torch.autograd.grad(want_out, xo, g)[0]
    xc = x.to(cuda).requires_grad_()
    with torch.enable_grad():
        out = edm.denoiser(xc, net, sigma.to(cuda))
        got = torch.autograd.grad(out, xc, g.to(cuda))[0]
    e_out, e_g = rel_l2(out, want_out), rel_l2(got, want)
    print(f"conv_mode {mode}: forward {e_out:.3e}, input gradient {e_g:.3e} vs oracle autograd")
    assert e_out < (1e-4 if mode == 0 else 1e-3)
    assert e_g < (1e-4 if mode == 0 else 2e-3)
    # bare module call (no preconditioning), per-clip sigma, and the no-grad path still works afterwards
    cn = torch.tensor([[-0.5], [0.1]])
    xo2 = x.clone().requires_grad_()
    w2 = torch.autograd.grad(orc.differentiable(xo2, cn), xo2, g)[0]
    xc2 = x.to(cuda).requires_grad_()
    g2 = torch.autograd.grad(net(xc2, cn.to(cuda)), xc2, g.to(cuda))[0]
    assert rel_l2(g2, w2) < (1e-4 if mode == 0 else 2e-3)
    with torch.no_grad():
        assert rel_l2(net(xc2, cn.to(cuda)), orc(x, cn)) < (1e-4 if mode == 0 else 1e-3)


def test_stale_tape_is_refused(aid, cuda):
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1))
    x = seeded((1, cfg.audio_len), 0).to(cuda).requires_grad_()
    cn = torch.tensor([[-0.5]], device=cuda)
    a = net(x, cn)
    b = net(x, cn)          # overwrites the tape of `a`
    torch.autograd.grad(b.sum(), x)
    with pytest.raises(RuntimeError, match="tape"):
        torch.autograd.grad(a.sum(), x)


@pytest.mark.parametrize("name,consistency", [("small_sample_guided_T6", True), ("small_sample_guided_noproj_T6", False)])
def test_guided_sampler_matches_reference_golden(aid, cuda, name, consistency):
    """The reference's default mode (xi = 0.25): 6 Heun steps = 11 denoiser evaluations, each with a backward pass, against
    the trajectory its own Sampler + unet.py produced on the CPU (batch 1)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_guided.npz"))
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1234))
    args = _tester_args(aid, T=6)
    args["tester"]["posterior_sampling"]["xi"] = 0.25
    args["tester"]["data_consistency"]["use"] = consistency
    args["exp"]["audio_len"] = cfg.audio_len
    s = aid.Sampler(net, aid.EDM(args), args)
    y = seeded((1, cfg.audio_len), 7, 0.063)
    mask = torch.ones(1, cfg.audio_len)
    mask[..., cfg.audio_len // 2 - 750: cfg.audio_len // 2 + 750] = 0
    torch.manual_seed(44)
    out = s.predict_inpainting((y * mask).to(cuda), mask.to(cuda))
    e = rel_l2(out, torch.from_numpy(g[name]))
    print(f"guided trajectory ({name}) vs the reference's run: {e:.3e}")
    assert e < 1e-3


def test_guided_sampler_batch_gt_1_is_per_clip(aid, cuda):
    """The reference crashes for batch > 1 on this branch; here clips are guided independently: a batch equals its clips run alone."""
    cfg = aid.small_test(16384)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(aid.random_state_dict(cfg, seed=1234))
    args = _tester_args(aid, T=3)
    args["tester"]["posterior_sampling"]["xi"] = 0.25
    args["exp"]["audio_len"] = cfg.audio_len
    s = aid.Sampler(net, aid.EDM(args), args)
    y = seeded((2, cfg.audio_len), 7, 0.063).to(cuda)
    mask = torch.ones(1, cfg.audio_len, device=cuda)
    mask[..., 8000:8600] = 0
    noise = [seeded((2, cfg.audio_len), 100 + k) for k in range(4)]
    s.noise_source = iter(noise)
    both = s.predict_inpainting(y * mask, mask)
    for k in range(2):
        s.noise_source = iter([n[k:k + 1] for n in noise])
        solo = s.predict_inpainting((y * mask)[k:k + 1], mask)
        assert rel_l2(both[k:k + 1], solo) < 1e-4


def test_paper_network_vjp_vs_oracle_autograd(aid, cuda):
    """The 186 M-parameter network (BASELINE config 1 shape, 1 x 65536) in conv_mode 2, where the data gradient of the dilated 5x3
    layers runs on tcgen05 with per-tensor power-of-two scaling: input gradient against autograd through the oracle on the host."""
    cfg = aid.paper_22k(65536, conv_mode=2)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, cuda)
    net.load_state_dict(sd)
    orc = make_oracle(cfg, sd)
    torch.set_num_threads(os.cpu_count() or 1)
    x, g, cn = seeded((1, 65536), 0, 0.4), seeded((1, 65536), 9), torch.tensor([[-0.5]])
    xo = x.clone().requires_grad_()
    yo = orc.differentiable(xo, cn)
    want = torch.autograd.grad(yo, xo, g)[0]
    xc = x.to(cuda).requires_grad_()
    yc = net(xc, cn.to(cuda))
    got = torch.autograd.grad(yc, xc, g.to(cuda))[0]
    e_f, e_g = rel_l2(yc, yo), rel_l2(got, want)
    print(f"paper network 1 x 65536, conv_mode 2: forward {e_f:.3e}, input gradient {e_g:.3e} vs oracle autograd")
    assert e_f < 1e-3 and e_g < 1e-3
