"""GPU parity of the conv_mode 2 operand layouts (conv_tc2.cu) against a plain numpy restatement of their definition:
activations [B][G][F+2PF][T+2][64] fp16 (x16) and weights [n-tile][kf][G][kt][Ntile][64] fp16 (x1024), both with the eight
16-byte chunks of every 128-byte row XOR-swizzled (by the flattened padded pixel index, resp. the cout row, mod 8)."""
import ctypes as C

import numpy as np
import pytest
import torch

from util import seeded
from test_gpu_ops import _lib

pytestmark = pytest.mark.gpu


def act_layout_ref(x, PF):
    B, Cc, Fd, T = x.shape
    G, Tp, rows = (Cc + 63) // 64, T + 2, Fd + 2 * PF
    out = np.zeros((B, G, rows, Tp, 8, 8), dtype=np.float16)   # [.., position chunk, 8 halves]
    xs = (x.numpy() * 16).astype(np.float16)
    for g in range(G):
        for ch in range(8):
            c0 = g * 64 + ch * 8
            if c0 >= Cc:
                continue
            blk = xs[:, c0:c0 + 8]                                  # [B, 8, F, T]
            for r in range(Fd):
                gp = (r + PF) * Tp + 1 + np.arange(T)               # flattened padded pixel index
                pos = ch ^ (gp & 7)
                out[:, g, r + PF, 1 + np.arange(T), pos, :] = np.transpose(blk[:, :, r, :], (0, 2, 1))
    return out.reshape(-1)


def weight_layout_ref(w):
    Co, Ci, KF, KT = w.shape
    G = (Ci + 63) // 64
    Nt = 128 if (Co == 256 and KF * KT > 1) else min(Co, 256)   # multi-tap 256-cout layers run as two 128-wide n-tiles
    ws = (w.numpy() * 1024).astype(np.float16)
    out = np.zeros((Co // Nt, KF, G, KT, Nt, 8, 8), dtype=np.float16)
    for n in range(Co):
        for ch in range((Ci + 7) // 8):
            g, cl = divmod(ch, 8)
            seg = ws[n, ch * 8:ch * 8 + 8]                          # [8, KF, KT]
            out[n // Nt, :, g, :, n % Nt, cl ^ (n & 7), :] = np.transpose(seg, (1, 2, 0))
    return out.reshape(-1)


@pytest.mark.parametrize("case", [(1, 16, 16, 3, 128, 1, 1, 0), (2, 64, 32, 5, 20, 5, 3, 4), (1, 96, 96, 4, 256, 5, 3, 0),
                                   (1, 320, 512, 1, 64, 1, 1, 3), (1, 128, 256, 6, 64, 5, 3, 7),
                                   (2, 80, 16, 7, 10, 1, 1, 2), (1, 64, 64, 40, 4, 5, 3, 9), (1, 16, 16, 5, 300, 1, 1, 0)])
def test_tc2_operand_layouts(cuda, case):
    B, Ci, Co, Fd, T, KF, KT, PF = case
    L = _lib()
    x = seeded((B, Ci, Fd, T), 5)
    w = seeded((Co, Ci, KF, KT), 6, 0.05)
    na, nw = C.c_uint64(), C.c_uint64()
    L.check(L.lib().aid_debug_tc2_operands(None, None, B, Ci, Co, Fd, T, KF, KT, PF, None, None, C.byref(na), C.byref(nw)))
    a_out = torch.full((na.value,), float("nan"), dtype=torch.float16, device=cuda)
    w_out = torch.full((nw.value,), float("nan"), dtype=torch.float16, device=cuda)
    xd, wd = x.to(cuda), w.to(cuda)
    L.check(L.lib().aid_debug_tc2_operands(L.ptr(xd), L.ptr(wd), B, Ci, Co, Fd, T, KF, KT, PF, L.ptr(a_out), L.ptr(w_out), None, None))
    a_ref, w_ref = act_layout_ref(x, PF), weight_layout_ref(w)
    assert a_ref.size == na.value and w_ref.size == nw.value
    a_got, w_got = a_out.cpu().numpy(), w_out.cpu().numpy()
    # channel positions past C in the last group are don't-care for the MMAs but are written as zeros
    assert np.array_equal(w_got, w_ref), f"weights differ at {np.flatnonzero(w_got != w_ref)[:8]}"
    assert np.array_equal(a_got, a_ref), f"activations differ at {np.flatnonzero(a_got != a_ref)[:8]} of {a_ref.size}"
