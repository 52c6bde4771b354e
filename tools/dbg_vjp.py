"""Bisects the denoiser VJP against oracle autograd over network variants: python tools/dbg_vjp.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch, aid_b200 as aid
from util import rel_l2, seeded, make_oracle
dev = torch.device("cuda:0")
L = 16384
variants = {
    "default": dict(),
    "no_attn": dict(attention_layers=[0] * 8),
    "no_attn_1dil": dict(attention_layers=[0] * 8, num_dils=[1] * 7),
    "attn_1dil": dict(num_dils=[1] * 7),
    "same_width": dict(Ns=[16] * 7, attention_layers=[0] * 8, num_dils=[1] * 7),
}
for name, kw in variants.items():
    base = dict(audio_len=L, Ns=[16, 16, 24, 24, 32, 32, 32], num_dils=[1, 2, 2, 3, 3, 3, 2], attention_layers=[0, 0, 0, 0, 1, 1, 1, 1])
    base.update(kw)
    cfg = aid.NetConfig(**base)
    sd = aid.random_state_dict(cfg, seed=1234)
    net = aid.Unet_CQT_oct_with_attention(cfg, dev); net.load_state_dict(sd)
    orc = make_oracle(cfg, sd)
    for B in (1, 2):
        x = seeded((B, L), 0, 0.4); g = seeded((B, L), 9); cn = torch.tensor([[-0.5]])
        xo = x.clone().requires_grad_(); yo = orc.differentiable(xo, cn); want = torch.autograd.grad(yo, xo, g)[0]
        xc = x.to(dev).requires_grad_(); yc = net(xc, cn.to(dev)); got = torch.autograd.grad(yc, xc, g.to(dev))[0].cpu()
        ratio = float((got * want).sum() / (want * want).sum())
        print(f"{name:14s} B={B}: fwd {rel_l2(yc, yo):.2e}  grad {rel_l2(got, want):.3e}  projection coefficient {ratio:.4f}  residual after scaling {rel_l2(got / ratio, want):.3e}", flush=True)
