"""GPU parity of the fused out block (out_block_kernel, out_block.cu; reference unet.py:452-493 as instantiated at unet.py:690, 719 and
called at unet.py:794, 817: one gated 1x1 residual layer N -> N, proj_out and res_conv N -> 2, optional accumulation into the running
2-channel output) through the C ABI.  The fused kernel folds P diag(gate) H into a 2 x N matrix per clip and evaluates the block in
fp32, so it sits at fp32 rounding from the fp64 definition, while the un-fused conv_mode 2 path carries the fp16 operand rounding of
its 1x1 layer; the two agree within that rounding."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as Fn

from util import rel_l2, seeded
from test_gpu_ops import _lib

pytestmark = pytest.mark.gpu

CASES = [
    # B, N, F, T, accum
    (2, 64, 8, 256, True),
    (1, 64, 64, 512, False),
    (3, 96, 5, 128, True),
    (2, 96, 16, 384, False),
    (1, 128, 7, 252, True),      # F * T % 4 == 0 is all the kernel needs
    (2, 128, 16, 128, True),
    (2, 256, 9, 64, False),      # bottleneck-like
    (5, 256, 40, 64, True),
    (5, 64, 40, 1024, True),
]
A = 0.7071067811865476


def _ref(x, wH, wP, wR, gamma, affine, gate, accum):
    """(skip, branch): the part of the output that does not pass the 1x1 layer, and the part that does (fp64)."""
    B, N, Fd, T = x.shape
    xd = x.double()
    std = xd.reshape(B, 8, -1).std(dim=2, unbiased=True).reshape(B, 8, 1, 1, 1)
    xn = (xd.reshape(B, 8, N // 8, Fd, T) / (std + 1e-7)).reshape(B, N, Fd, T)
    a = Fn.gelu(xn * (gamma.double() * (1 + affine.double())).reshape(1, N, 1, 1))
    h = torch.einsum("nm,bmft->bnft", wH.double(), a) * gate.double().reshape(1, N, 1, 1)
    proj = lambda w, v: torch.einsum("kn,bnft->bkft", w.double(), v)
    skip = A * (A * proj(wP, xd) + proj(wR, xd))
    branch = A * A * proj(wP, h)
    if accum is not None:
        skip = A * (accum.double() + skip)
        branch = A * branch
    return skip, branch


def _inputs(case):
    B, N, Fd, T, acc = case
    x = seeded((B, N, Fd, T), 21)
    if B > 1: x[1] *= 2.5                                     # per-clip statistics
    wH = seeded((N, N), 22, 1.0 / math.sqrt(N))
    wP, wR = seeded((2, N), 23, 1.0 / math.sqrt(N)), seeded((2, N), 24, 1.0 / math.sqrt(N))
    gamma, affine, gate = 1 + 0.2 * seeded((N,), 25), 0.3 * seeded((N,), 26), seeded((N,), 27)
    accum = seeded((B, 2, Fd, T), 28) if acc else None
    return x, wH, wP, wR, gamma, affine, gate, accum


def _exec(cuda, ins, fused, time=False, alias=False):
    L = _lib()
    x, wH, wP, wR, gamma, affine, gate, accum = ins
    B, N, Fd, T = x.shape
    dev = [t.contiguous().to(cuda) for t in (x, wH, wP, wR, gamma, affine, gate)]     # keep the device copies alive
    acc_d = accum.to(cuda) if accum is not None else None
    out = acc_d if (alias and acc_d is not None) else torch.full((B, 2, Fd, T), float("nan"), device=cuda)
    ms = C.c_float()
    L.check(L.lib().aid_debug_out_block(L.ptr(dev[0]), L.ptr(dev[1]), L.ptr(dev[2]), L.ptr(dev[3]), B, N, Fd, T, L.ptr(dev[4]), L.ptr(dev[5]),
                                        L.ptr(dev[6]), L.ptr(acc_d) if acc_d is not None else None, fused, L.ptr(out),
                                        C.byref(ms) if time else None))
    torch.cuda.synchronize()
    return out, ms.value


@pytest.mark.parametrize("case", CASES)
def test_out_block_fused(cuda, case):
    ins = _inputs(case)
    o1, _ = _exec(cuda, ins, 1)
    o0, _ = _exec(cuda, ins, 0)
    assert torch.isfinite(o1).all()
    skip, branch = _ref(*ins)
    e1 = rel_l2(o1.cpu().double(), skip + branch)
    eb1 = rel_l2(o1.cpu().double() - skip, branch)
    eb0 = rel_l2(o0.cpu().double() - skip, branch)
    e01 = rel_l2(o1, o0)
    print(f"{case}: fused vs fp64 {e1:.2e} (layer branch {eb1:.2e}); un-fused layer branch vs fp64 {eb0:.2e}; fused vs un-fused {e01:.2e}")
    assert e1 < 2e-6 and eb1 < 1e-5            # fp32 arithmetic; the GELU is the 5e-7-absolute erf approximation of the operand pass
    assert eb0 < 1e-3                          # the conv_mode 2 bar of the un-fused layer (fp16 operands)
    assert e01 < 5e-4
    if ins[-1] is not None:                    # in the network the running output is updated in place
        o2, _ = _exec(cuda, ins, 1, alias=True)
        assert torch.equal(o2, o1)


def test_out_block_batch_invariant(cuda):
    ins = _inputs((5, 96, 40, 256, True))
    a, _ = _exec(cuda, ins, 1)
    b, _ = _exec(cuda, ins, 1)
    assert torch.equal(a, b)
    for k in (0, 1, 4):
        solo, _ = _exec(cuda, (ins[0][k:k + 1],) + ins[1:7] + (ins[7][k:k + 1],), 1)
        assert torch.equal(solo[0], a[k])


def test_out_block_timing(cuda):
    """Report (not assert) the out blocks of three levels at 8 clips of the bench shape."""
    for case in [(8, 64, 64, 4096, False), (8, 96, 192, 1024, False), (8, 128, 320, 256, False), (8, 256, 448, 64, False)]:
        ins = _inputs(case)
        _, t1 = _exec(cuda, ins, 1, time=True)
        _, t0 = _exec(cuda, ins, 0, time=True)
        B, N, Fd, T, _ = case
        gb = 4.0 * B * Fd * T * (2 + N) / 1e9
        print(f"{case}: fused {t1:.3f} ms ({gb / t1 * 1e3:.0f} GB/s algorithmic), un-fused {t0:.3f} ms")
