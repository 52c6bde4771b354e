"""GPU parity of the fused dilated residual layer (conv_comb_kernel: group-norm apply, adaLN modulation, GELU and the fp16 operand
conversion inside the tcgen05 convolution; reference unet.py:470-482) through the C ABI.  It must reproduce the two-kernel path
(operand pass + conv_tc2_kernel) bit for bit -- same operand roundings, same MMA order per accumulator -- and the fp64 definition
of the layer within the conv_mode 2 bar."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as Fn

from util import rel_l2, seeded
from test_gpu_ops import _lib

pytestmark = pytest.mark.gpu

CASES = [
    # B, F, T, dil[, C]
    (2, 16, 256, 1),
    (1, 13, 128, 2),      # odd F: combs of 7 and 6 rows
    (1, 5, 384, 4),       # combs of 2, 1, 1, 1 rows (shorter than the tap span)
    (3, 8, 128, 1),
    (1, 64, 512, 2),      # level-0-like
    (1, 3, 128, 4),       # dil > F: one comb is empty
    (2, 40, 256, 8),
    (1, 16, 256, 1, 96),  # 96 channels: 64-channel group (SWIZZLE_128B) + 32-channel group (SWIZZLE_64B), streamed weights, row pairs
    (2, 13, 128, 2, 96),
    (1, 5, 384, 4, 96),
    (1, 64, 512, 4, 96),
    (2, 3, 128, 4, 96),
    (5, 16, 1024, 4, 64),  # 160 combs: CTAs walk several items (clip changes, ring and accumulator phases carried across items)
    (5, 16, 1024, 4, 96),
    (3, 8, 1024, 8, 96),   # 192 one-row combs
]


def _layer_ref(x, w, gamma, affine, gate, alpha, dil):
    B, Cn, Fd, T = x.shape
    xd = x.double()
    g = xd.reshape(B, 8, -1)
    std = g.std(dim=2, unbiased=True).reshape(B, 8, 1, 1, 1)
    xn = (xd.reshape(B, 8, Cn // 8, Fd, T) / (std + 1e-7)).reshape(B, Cn, Fd, T)
    a = Fn.gelu(xn * (gamma.double() * (1 + affine.double())).reshape(1, Cn, 1, 1))
    y = Fn.conv2d(a, w.double(), padding=(2 * dil, 1), dilation=(dil, 1))
    return alpha * (xd + y * gate.double().reshape(1, Cn, 1, 1))


def _run(cuda, case, fused, time=False):
    B, Fd, T, dil = case[:4]
    Cn = case[4] if len(case) > 4 else 64
    L = _lib()
    x = seeded((B, Cn, Fd, T), 1)
    w = seeded((Cn, Cn, 5, 3), 2, 1.0 / math.sqrt(Cn * 15))
    gamma, affine, gate = 1 + 0.2 * seeded((Cn,), 3), 0.3 * seeded((Cn,), 4), seeded((Cn,), 5)
    out = torch.full((B, Cn, Fd, T), float("nan"), device=cuda)
    stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=cuda)
    ms = C.c_float()
    xd, wd, gd, ad, gtd = x.to(cuda), w.to(cuda), gamma.to(cuda), affine.to(cuda), gate.to(cuda)      # keep the device copies alive
    L.check(L.lib().aid_debug_dilated_layer(L.ptr(xd), L.ptr(wd), B, Cn, Fd, T, dil, L.ptr(gd), L.ptr(ad), L.ptr(gtd), 0.70710678, fused,
                                            L.ptr(out), L.ptr(stats), C.byref(ms) if time else None))
    torch.cuda.synchronize()
    return out, stats, (x, w, gamma, affine, gate), ms.value


@pytest.mark.parametrize("case", CASES)
def test_conv_comb_equals_two_kernel_path(cuda, case):
    o1, s1, ins, _ = _run(cuda, case, 1)
    o0, s0, _, _ = _run(cuda, case, 0)
    assert torch.isfinite(o1).all()
    assert torch.equal(o0, o1)          # same operand roundings, same MMA order per accumulator (kf, channel group, kt, k-step)
    assert torch.allclose(s0, s1, rtol=1e-9, atol=1e-6)
    x, w, gamma, affine, gate = ins
    ref = _layer_ref(x, w, gamma, affine, gate, 0.70710678, case[3])
    assert rel_l2(o1.cpu().double() - 0.70710678 * x.double(), ref - 0.70710678 * x.double()) < 1e-3
    # the statistics describe the output the kernel wrote (fp32 partial sums accumulated in double)
    g = o1.cpu().double().reshape(case[0], 8, -1)
    assert torch.allclose(s1[:, :, 0].cpu(), g.sum(-1), rtol=1e-6, atol=1e-6 * float(g.abs().sum(-1).max()))
    assert torch.allclose(s1[:, :, 1].cpu(), (g * g).sum(-1), rtol=1e-5, atol=0)
