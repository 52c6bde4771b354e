#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vjp.py -m gpu -q -s -k "denoiser_vjp or guided or stale" > gpurun_out/r2_t14.log 2>&1
echo "vjp net tests rc=$?"; grep -v "^$" gpurun_out/r2_t14.log | grep "conv_mode\|guided\|passed\|failed\|^E " | head -20
python - <<'PY'
import sys, os
sys.path[:0] = [".", "oracle", "tests"]
import torch, aid_b200 as aid
from util import rel_l2, seeded, make_oracle
# gradient parity at the paper network, 1 x 65536 (oracle autograd on the host), conv_mode 2 with the tensor-core data gradient
dev = torch.device("cuda:0")
cfg = aid.paper_22k(65536, conv_mode=2)
sd = aid.random_state_dict(cfg, seed=1234)
net = aid.Unet_CQT_oct_with_attention(cfg, dev); net.load_state_dict(sd)
orc = make_oracle(cfg, sd)
torch.set_num_threads(os.cpu_count())
x = seeded((1, 65536), 0, 0.4); g = seeded((1, 65536), 9); cn = torch.tensor([[-0.5]])
xo = x.clone().requires_grad_(); yo = orc.differentiable(xo, cn); want = torch.autograd.grad(yo, xo, g)[0]
xc = x.to(dev).requires_grad_(); yc = net(xc, cn.to(dev)); got = torch.autograd.grad(yc, xc, g.to(dev))[0]
print(f"paper network 1 x 65536 conv_mode 2: forward {rel_l2(yc, yo):.3e}, input gradient {rel_l2(got, want):.3e}")
PY
python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 35 --xi 0.25 > gpurun_out/r2_samp_guided_b1.json 2>> gpurun_out/r2_vjp.err
cat gpurun_out/r2_samp_guided_b1.json
