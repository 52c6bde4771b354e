"""One denoiser forward bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_forward.py --batch 4
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_simt -c 3 \
        -o gpurun_out/conv python tools/profile_forward.py --batch 4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aid_b200

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--len", type=int, default=262144)
ap.add_argument("--conv-mode", type=int, default=2)
ap.add_argument("--warm", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = aid_b200.paper_22k(a.len, conv_mode=a.conv_mode)
net = aid_b200.Unet_CQT_oct_with_attention(cfg, dev)
net.load_state_dict(aid_b200.random_state_dict(cfg, seed=1234))
x = torch.randn(a.batch, a.len, device=dev) * 0.5
cn = torch.tensor([[-0.3]], device=dev)
for _ in range(a.warm):
    net(x, cn)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net(x, cn)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
