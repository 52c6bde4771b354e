"""Precision study (CPU, oracle only): whole-network rel-L2 error of the denoiser when the operands of the dense
convolutions (the layers the tcgen05 kernel takes) are rounded the way a candidate tensor-core scheme would round them.
Decides how many MMAs per algorithmic MAC the GPU path must spend to stay inside the 1e-3 parity bar.

    python tools/precision_study.py [audio_len] [test_mode 0|1]

Schemes (a = activation operand, w = weight; h() = round to fp16, b() = bf16, q() = fp8 e4m3 with a power-of-two scale):
    f16x1   h(a)*h(w)                                      1 MMA
    f16x2a  (h(a)+h(a-h(a)))*h(w)                          2 MMAs   (activation split, weight single)
    f16x2w  h(a)*(h(w)+h(w-h(w)))                          2 MMAs
    f16x3   a_hi*w_hi + a_lo*w_hi + a_hi*w_lo              3 MMAs   (current conv_mode 1)
    f16+f8  a_hi*w_hi + q(a_lo)*q(w) + q(a)*q(w_lo)        1 fp16 + 2 fp8 MMAs = 2 fp16-equivalents
    bf16x1  b(a)*b(w)
A second table measures the GELU of the operand pass: f16x1 with the exact erf replaced by the Abramowitz-Stegun rational
approximations 7.1.26 (5 coefficients, what gn_act_tc2_kernel evaluates) and 7.1.25 (3 coefficients, 2 FMAs fewer per element).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import aid_b200  # noqa: E402
import unet_oracle  # noqa: E402
from util import make_oracle, rel_l2  # noqa: E402

A_SCALE, W_SCALE = 16.0, 1024.0


def h(x):
    return x.clamp(-60000, 60000).half().float()


def b(x):
    return x.bfloat16().float()


def q8(x, scale):
    return (x * scale).clamp(-448, 448).to(torch.float8_e4m3fn).float() / scale


def make_conv(scheme):
    def conv(sd, name, x, dilation=1):
        w = sd[name + ".weight"]
        dense = w.shape[1] % 16 == 0 and w.shape[0] % 16 == 0 and w.dim() == 4
        if scheme == "fp32" or not dense:
            return F.conv2d(x, w, padding="same", dilation=dilation)
        a = x * A_SCALE
        ww = w * W_SCALE
        cv = lambda aa, bb: F.conv2d(aa, bb, padding="same", dilation=dilation)
        if scheme == "f16x1":
            y = cv(h(a), h(ww))
        elif scheme == "bf16x1":
            y = cv(b(a), b(ww))
        elif scheme == "f16x2a":
            ah = h(a); al = h(a - ah)
            y = cv(ah + al, h(ww))
        elif scheme == "f16x2w":
            wh = h(ww); wl = h(ww - wh)
            y = cv(h(a), wh + wl)
        elif scheme == "f16x3":
            ah = h(a); al = h(a - ah); wh = h(ww); wl = h(ww - wh)
            y = cv(ah, wh) + cv(al, wh) + cv(ah, wl)
        elif scheme == "f16+f8":
            ah = h(a); al = a - ah; wh = h(ww); wl = ww - wh
            # fp8 operands: a (scale 1), w (scale 2^2), lo parts scaled by 2^11
            y = cv(ah, wh) + cv(q8(al, 2048.0), q8(ww, 4.0)) + cv(q8(a, 1.0), q8(wl, 4.0 * 2048.0))
        else:
            raise ValueError(scheme)
        return y / (A_SCALE * W_SCALE)
    return conv


def gelu_as(order):
    """x * Phi(x) with erf from A&S 7.1.26 (order 5) or 7.1.25 (order 3), evaluated in fp32 like the kernel does."""
    def g(v):
        ax = v.abs() * 0.7071067811865476
        if order == 5:
            t = 1.0 / (1.0 + 0.3275911 * ax)
            pl = ((((1.061405429 * t - 1.453152027) * t + 1.421413741) * t - 0.284496736) * t + 0.254829592) * t
        else:
            t = 1.0 / (1.0 + 0.47047 * ax)
            pl = ((0.7478556 * t - 0.0958798) * t + 0.3480242) * t
        erf_abs = 1.0 - pl * torch.exp(-ax * ax)
        return 0.5 * v * (1.0 + torch.sign(v) * erf_abs)
    return g


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    test_mode = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    torch.set_num_threads(os.cpu_count())
    cfg = aid_b200.paper_22k(L)
    sd = aid_b200.random_state_dict(cfg, seed=1234, test_mode=test_mode)
    orc = make_oracle(cfg, sd)
    orig = unet_oracle.conv
    for sigma in (1.0, 0.05):
        x = torch.randn(1, L, generator=torch.Generator().manual_seed(0)) * (sigma ** 2 + 0.063 ** 2) ** 0.5
        cin = (0.063 ** 2 + sigma ** 2) ** -0.5
        cn = torch.tensor([[0.25 * torch.log(torch.tensor(sigma)).item()]])
        unet_oracle.conv = orig
        ref = orc(cin * x, cn)
        for scheme in ("f16x1", "bf16x1", "f16x2a", "f16x2w", "f16+f8", "f16x3"):
            unet_oracle.conv = make_conv(scheme)
            out = orc(cin * x, cn)
            print(f"L={L} test_mode={int(test_mode)} sigma={sigma:<5} {scheme:7s} rel-L2 = {rel_l2(out, ref):.3e}", flush=True)
        orig_gelu = F.gelu
        try:
            for scheme in ("fp32", "f16x1"):       # the approximation alone, then together with the operand rounding
                unet_oracle.conv = make_conv(scheme)
                for name, order in (("A&S 7.1.26", 5), ("A&S 7.1.25", 3)):
                    unet_oracle.F.gelu = lambda v, _g=gelu_as(order): _g(v)
                    out = orc(cin * x, cn)
                    print(f"L={L} test_mode={int(test_mode)} sigma={sigma:<5} {scheme:7s} + GELU {name} rel-L2 = {rel_l2(out, ref):.3e}", flush=True)
        finally:
            unet_oracle.F.gelu = orig_gelu
    unet_oracle.conv = orig


if __name__ == "__main__":
    main()
