"""GPU parity of every single operator of the denoiser path against plain PyTorch fp32 (CPU) through the C ABI.

Tolerances are fp32 round-off of differently ordered sums: rel-L2 <= 2e-6 unless stated.
"""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from util import rel_l2, seeded

pytestmark = pytest.mark.gpu


def _lib():
    import aid_b200
    from aid_b200 import _lib
    return _lib


def conv_ref(a, w, dil, gate=None, R=None, R2=None, alpha=1.0, beta=0.0):
    y = F.conv2d(a.double(), w.double(), padding="same", dilation=(dil, 1))
    if gate is not None:
        y = y * gate.double().view(1, -1, 1, 1)
    if R is not None:
        y = y + R.double()
    y = alpha * y
    if R2 is not None:
        y = y + beta * R2.double()
    return y


CONV_CASES = [
    # B, Cin, Cout, F, T, KF, KT, dil, gate, R, R2, stats
    (2, 64, 64, 64, 256, 5, 3, 1, True, True, False, True),
    (1, 64, 64, 64, 128, 5, 3, 2, True, True, False, True),
    (1, 32, 32, 100, 16, 5, 3, 16, True, True, False, True),    # level-6-like: most taps in the zero padding
    (1, 96, 96, 40, 8, 5, 3, 64, True, True, False, True),      # dilation larger than F
    (1, 24, 24, 33, 20, 5, 3, 4, True, True, False, True),      # ragged F/T, group size 3
    (2, 2, 64, 64, 64, 5, 3, 1, False, True, False, True),      # pyr_down_proj
    (2, 2, 96, 64, 64, 1, 1, 1, False, False, False, True),     # init-block proj_in
    (1, 128, 96, 17, 50, 1, 1, 1, False, True, False, True),    # res_conv with residual
    (1, 64, 2, 30, 64, 1, 1, 1, False, True, True, False),      # out block: alpha*(acc+R)+beta*R2
    (1, 96, 8, 12, 40, 1, 1, 1, False, False, False, False),    # attention proj_in
    (1, 8, 96, 12, 40, 1, 1, 1, True, True, False, True),       # attention proj_out
    (1, 200, 400, 1, 32, 1, 1, 1, False, False, False, False),  # qk Conv1d as [B, 8F, 1, T]
    (1, 16, 16, 9, 4, 5, 3, 2, True, True, False, True),        # T < 8
]


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("case", CONV_CASES + [
    (2, 2, 64, 9, 64, 1, 1, 1, False, True, False, True),      # thin-in 1x1 with residual + statistics (init res_conv)
    (1, 8, 24, 5, 16, 1, 1, 1, True, True, False, True),       # thin-in 8 -> N, groups of 3 channels
    (1, 2, 40, 20, 6, 5, 3, 1, False, True, False, True),      # thin-in 5x3, T = 6 (pair-vectorised)
    (2, 48, 8, 7, 12, 1, 1, 1, False, False, False, False),    # thin-out N -> 8
    (1, 24, 2, 6, 8, 1, 1, 1, False, True, True, False),       # thin-out N -> 2 with both residuals
    (2, 64, 2, 30, 64, 5, 3, 1, False, False, True, False),    # thin-out 5x3 N -> 2 accumulating into R2 (pyramid projection gradient)
    (1, 40, 2, 7, 8, 5, 3, 1, True, True, True, False),        # ... gate and both residuals, F of the order of the tap span
])
def test_conv2d(cuda, case, mode):
    B, Cin, Cout, Fd, T, KF, KT, dil, use_gate, use_R, use_R2, use_stats = case
    L = _lib()
    a = seeded((B, Cin, Fd, T), 1)
    w = seeded((Cout, Cin, KF, KT), 2, 1.0 / math.sqrt(Cin * KF * KT))
    gate = seeded((Cout,), 3) if use_gate else None
    R = seeded((B, Cout, Fd, T), 4) if use_R else None
    R2 = seeded((B, Cout, Fd, T), 5) if use_R2 else None
    alpha, beta = 0.70710678, 0.5
    ref = conv_ref(a, w, dil, gate, R, R2, alpha, beta)
    d = lambda t: None if t is None else t.to(cuda).contiguous()
    ad, wd, gd, Rd, R2d = d(a), d(w), d(gate), d(R), d(R2)
    out = torch.empty(B, Cout, Fd, T, device=cuda)
    stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=cuda) if use_stats else None
    L.check(L.lib().aid_op_conv2d(L.ptr(ad), L.ptr(wd), B, Cin, Cout, Fd, T, KF, KT, dil, L.ptr(gd), L.ptr(Rd), L.ptr(R2d),
                                  alpha, beta, L.ptr(out), L.ptr(stats), mode, None))
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-6
    if use_stats:
        g = ref.reshape(B, 8, -1)
        assert torch.allclose(stats[:, :, 0].cpu(), g.sum(-1), rtol=1e-5, atol=1e-4)
        assert torch.allclose(stats[:, :, 1].cpu(), (g * g).sum(-1), rtol=1e-5, atol=1e-4)


def test_conv2d_inplace_residual(cuda):
    """x' = (x + conv(a)*g)/sqrt2 written over x (how the residual layers run, unet.py:482)."""
    L = _lib()
    B, Cn, Fd, T = 1, 32, 20, 64
    a, x, w, g = seeded((B, Cn, Fd, T), 1), seeded((B, Cn, Fd, T), 2), seeded((Cn, Cn, 5, 3), 3, 0.05), seeded((Cn,), 4)
    ref = conv_ref(a, w, 2, g, x, None, 2 ** -0.5)
    ad, xd, wd, gd = a.to(cuda), x.to(cuda), w.to(cuda), g.to(cuda)
    L.check(L.lib().aid_op_conv2d(L.ptr(ad), L.ptr(wd), B, Cn, Cn, Fd, T, 5, 3, 2, L.ptr(gd), L.ptr(xd), None,
                                  2 ** -0.5, 0.0, L.ptr(xd), None, 0, None))
    torch.cuda.synchronize()
    assert rel_l2(xd, ref) < 2e-6


@pytest.mark.parametrize("shape,gelu", [((2, 64, 64, 128), True), ((1, 96, 33, 20), True), ((2, 16, 7, 4), False), ((1, 256, 448, 16), True)])
def test_groupnorm_act(cuda, shape, gelu):
    L = _lib()
    B, Cn, Fd, T = shape
    x = seeded(shape, 1, 3.0) + 0.5  # not centred: the reference does not subtract the mean (unet.py:155-159)
    gamma, aff = 1 + 0.3 * seeded((Cn,), 2), 0.5 * seeded((Cn,), 3)
    xg = x.double().reshape(B, 8, -1)
    ref = (xg / (xg.std(-1, keepdim=True) + 1e-7)).reshape(shape) * gamma.double().view(1, -1, 1, 1) * (aff.double().view(1, -1, 1, 1) + 1)
    if gelu:
        ref = F.gelu(ref)
    xd, gd, ad = x.to(cuda), gamma.to(cuda), aff.to(cuda)
    out = torch.empty_like(xd)
    scratch = torch.empty(B * 16, dtype=torch.float64, device=cuda)
    L.check(L.lib().aid_op_groupnorm_act(L.ptr(xd), L.ptr(gd), L.ptr(ad), B, Cn, Fd, T, int(gelu), L.ptr(out), L.ptr(scratch), None))
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-6


@pytest.mark.parametrize("shape", [(2, 3, 5, 64), (1, 2, 64, 4), (1, 4, 3, 1024), (1, 1, 2, 6)])
@pytest.mark.parametrize("up", [0, 1])
def test_resample(cuda, shape, up):
    import unet_oracle
    L = _lib()
    B, Cn, Fd, T = shape
    x = seeded(shape, 7)
    ref = unet_oracle.up_t(x) if up else unet_oracle.down_t(x)
    xd = x.to(cuda)
    out = torch.empty(B, Cn, Fd, 2 * T if up else T // 2, device=cuda)
    L.check(L.lib().aid_op_resample(L.ptr(xd), B, Cn, Fd, T, up, L.ptr(out), None))
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < 1e-6


def test_resample_impulse_and_dc(cuda):
    """The facts SURVEY.md App. A records for the reference resampler: down has DC gain 1, up has DC gain 0.5."""
    L = _lib()
    ones = torch.ones(1, 1, 1, 32, device=cuda)
    dn, up = torch.empty(1, 1, 1, 16, device=cuda), torch.empty(1, 1, 1, 64, device=cuda)
    L.check(L.lib().aid_op_resample(L.ptr(ones), 1, 1, 1, 32, 0, L.ptr(dn), None))
    L.check(L.lib().aid_op_resample(L.ptr(ones), 1, 1, 1, 32, 1, L.ptr(up), None))
    torch.cuda.synchronize()
    assert torch.allclose(dn, torch.ones_like(dn), atol=1e-6)
    assert torch.allclose(up, 0.5 * torch.ones_like(up), atol=1e-6)


@pytest.mark.parametrize("B,heads,Fd,T", [(2, 8, 64, 16), (1, 8, 320, 64), (1, 8, 448, 256), (1, 4, 40, 50), (1, 8, 33, 300)])
def test_attention(cuda, B, heads, Fd, T):
    L = _lib()
    h = seeded((B, heads, Fd, T), 1)
    qk = seeded((B, 2 * heads * Fd, T), 2, 1.5 / math.sqrt(math.sqrt(Fd)))
    q4 = qk.double().reshape(B, heads, 2 * Fd, T).permute(0, 1, 3, 2)
    q, k = q4[..., :Fd], q4[..., Fd:]
    sim = torch.matmul(q, k.transpose(-1, -2)) * Fd ** -0.5
    ref = torch.matmul(sim.softmax(-1), h.double().permute(0, 1, 3, 2)).permute(0, 1, 3, 2)
    hd, qd = h.to(cuda), qk.to(cuda)
    out = torch.empty_like(hd)
    L.check(L.lib().aid_op_attention(L.ptr(hd), L.ptr(qd), B, heads, Fd, T, L.ptr(out), None))
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 3e-6


@pytest.mark.parametrize("B,heads,Fd,T", [(2, 8, 64, 16), (1, 8, 320, 256), (2, 8, 384, 128), (1, 8, 448, 64), (1, 8, 448, 256), (1, 4, 48, 50),
                                          (1, 8, 512, 32), (3, 2, 80, 200)])
def test_attention_tcgen05(cuda, B, heads, Fd, T):
    """The tcgen05 attention core of conv_mode 2 (Q K^T split fp16 -> TMEM, softmax from TMEM, P V fp16) against the fp64
    definition (unet.py:353-374): the published bar is 1e-3; the logits are fp32-grade, P and V carry fp16 rounding."""
    L = _lib()
    h = seeded((B, heads, Fd, T), 1)
    qk = seeded((B, 2 * heads * Fd, T), 2, 1.5 / math.sqrt(math.sqrt(Fd)))
    q4 = qk.double().reshape(B, heads, 2 * Fd, T).permute(0, 1, 3, 2)
    q, k = q4[..., :Fd], q4[..., Fd:]
    sim = torch.matmul(q, k.transpose(-1, -2)) * Fd ** -0.5
    ref = torch.matmul(sim.softmax(-1), h.double().permute(0, 1, 3, 2)).permute(0, 1, 3, 2)
    hd, qd = h.to(cuda), qk.to(cuda)
    out = torch.full_like(hd, float("nan"))
    L.check(L.lib().aid_op_attention_mode(L.ptr(hd), L.ptr(qd), B, heads, Fd, T, L.ptr(out), 1, None))
    torch.cuda.synchronize()
    e = rel_l2(out, ref)
    print(f"tcgen05 attention B{B} heads{heads} F{Fd} T{T}: rel-L2 {e:.2e}")
    assert torch.isfinite(out).all() and e < 5e-4
    # sharper logits (larger q, k): the softmax is close to one-hot, the split operands keep the logits accurate
    qk2 = (qk * 4).to(cuda)
    sim2 = sim * 16
    ref2 = torch.matmul(sim2.softmax(-1), h.double().permute(0, 1, 3, 2)).permute(0, 1, 3, 2)
    L.check(L.lib().aid_op_attention_mode(L.ptr(hd), L.ptr(qk2), B, heads, Fd, T, L.ptr(out), 1, None))
    assert rel_l2(out, ref2) < 5e-4
