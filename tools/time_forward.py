"""ms per denoiser forward (paper network, 262144 samples) at small batches: python tools/time_forward.py [batches ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, aid_b200
dev = torch.device("cuda:0")
cfg = aid_b200.paper_22k(262144, conv_mode=2)
net = aid_b200.Unet_CQT_oct_with_attention(cfg, dev)
net.load_state_dict(aid_b200.random_state_dict(cfg, seed=1234))
cn = torch.tensor([[-0.3]], device=dev)
for B in [int(v) for v in sys.argv[1:]] or [1, 2, 4, 8]:
    x = torch.randn(B, 262144, device=dev) * 0.5
    for _ in range(3): net(x, cn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): net(x, cn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"B={B}: {ms:.2f} ms per forward, {B / ms * 1e3:.1f} clips/s", flush=True)
