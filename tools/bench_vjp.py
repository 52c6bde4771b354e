"""Times the differentiable denoiser call of the guidance branch (sampler.py:57-113): plain forward, taped forward, backward.

    python tools/bench_vjp.py [--len 262144] [--batch 1] [--conv-mode 2] [--net paper_22k|paper_44k]
One JSON line: ms per plain forward, taped forward and backward (CUDA events, median of 5 after 2 warm-ups), workspace bytes.
The reference's only published timing is this mode: 0.26 s per denoiser forward + backward, MusicNet 44.1 kHz 8-octave network,
L = 184184, batch 1, on an A100 (notebooks/demo_inpainting_spectrogram.ipynb cell 8)."""
import argparse, ctypes as C, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, aid_b200
from aid_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--len", type=int, default=262144)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--conv-mode", type=int, default=2)
ap.add_argument("--net", default="paper_22k", choices=["paper_22k", "paper_44k"])
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = getattr(aid_b200, a.net)(a.len, conv_mode=a.conv_mode)
net = aid_b200.Unet_CQT_oct_with_attention(cfg, dev)
net.load_state_dict(aid_b200.random_state_dict(cfg, seed=1234))
x = (torch.randn(a.batch, a.len, generator=torch.Generator().manual_seed(0)) * 0.5).to(dev)
g = torch.randn(a.batch, a.len, generator=torch.Generator().manual_seed(1)).to(dev)
cn = torch.tensor([[-0.3]], device=dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


with torch.no_grad():
    ms_fwd = timed(lambda: net(x, cn))
ms_tape = timed(lambda: net._forward_tape(x, cn, 1.0, 1.0, 0.0))
net._forward_tape(x, cn, 1.0, 1.0, 0.0)
ms_bwd = timed(lambda: net._backward(g))
need = C.c_size_t()
_lib.check(_lib.lib().aid_vjp_workspace_bytes(net._handle, a.batch, C.byref(need)), net._handle)
print(json.dumps({"net": a.net, "len": a.len, "batch": a.batch, "conv_mode": a.conv_mode, "ms_forward": ms_fwd, "ms_forward_taped": ms_tape,
                  "ms_backward": ms_bwd, "ms_forward_plus_backward": ms_tape + ms_bwd, "vjp_workspace_gb": need.value / 1e9,
                  "reference_published": "0.26 s per forward + backward (A100, 44.1 kHz 8-octave network, L = 184184, batch 1)"}))
