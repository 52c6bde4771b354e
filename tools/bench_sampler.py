"""End-to-end sampling throughput (BASELINE configs 2-5) on the B200 path.

    python tools/bench_sampler.py --config inpaint --batch 32 --gap-ms 300      # config 3
    python tools/bench_sampler.py --config uncond  --batch 8                    # config 2
    torchrun --nproc-per-node N ... tools/bench_sampler.py --config inpaint --batch 256 --gap-ms 1500   # config 4 (sharded)
Prints one JSON line (rank 0): clips/s for the whole 35-step, 69-evaluation run, seconds per run, evals.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import aid_b200
from aid_b200.dist import ShardedSampler, shard_bounds

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="inpaint", choices=["inpaint", "uncond"])
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--len", type=int, default=262144)
ap.add_argument("--gap-ms", type=float, default=300.0)
ap.add_argument("--steps", type=int, default=35)
ap.add_argument("--conv-mode", type=int, default=2)
ap.add_argument("--noise", default="device", choices=["device", "host"], help="device: Philox on the GPU + CUDA-graph replay; host: per-clip torch generators + H2D")
ap.add_argument("--no-graph", action="store_true", help="device noise, eager launches")
ap.add_argument("--gap-sweep", action="store_true", help="BASELINE config 5: one timed run per gap length of the reference's sweep")
ap.add_argument("--xi", type=float, default=0.0, help="> 0: reconstruction guidance (the reference's default 0.25), host noise, per-clip")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

cfg = aid_b200.paper_22k(a.len, conv_mode=a.conv_mode)
net = aid_b200.Unet_CQT_oct_with_attention(cfg, dev)
net.load_state_dict(aid_b200.random_state_dict(cfg, seed=1234))
args = aid_b200.AttrDict.wrap({
    "tester": {"T": a.steps, "order": 2, "filter_out_cqt_DC_Nyq": True, "posterior_sampling": {"xi": a.xi, "norm": 2, "smoothl1_beta": 1},
               "data_consistency": {"use": True, "type": "always", "smooth": True, "hann_size": 50},
               "diff_params": {"same_as_training": False, "sigma_data": 0.063, "sigma_min": 1e-4, "sigma_max": 1, "ro": 13,
                               "Schurn": 10, "Snoise": 1.0, "Stmin": 0, "Stmax": 50}},
    "diff_params": {"sigma_data": 0.063, "sigma_min": 1e-5, "sigma_max": 10, "ro": 13, "Schurn": 5, "Snoise": 1, "Stmin": 0, "Stmax": 50},
    "exp": {"audio_len": a.len, "sample_rate": 22050},
})
if a.xi > 0:
    a.noise = "host"          # the guidance branch runs the reference-ordered torch loop around the differentiable CUDA denoiser
smp = ShardedSampler(aid_b200.Sampler(net, aid_b200.EDM(args), args), seed=42, device_noise=(a.noise == "device"))
smp.sampler.use_cuda_graph = not a.no_graph
B, L = a.batch, a.len
# MAESTRO-shaped synthetic clips: zero mean, RMS = sigma_data
y = (torch.randn(B, L, generator=torch.Generator().manual_seed(7)) * 0.063).to(dev)
GAPS_MS = [25, 50, 100, 300, 371, 743, 1486, 1500] if a.gap_sweep else [a.gap_ms]   # inpainting_tester.yaml:71,75; tester_inpainting.py:355-357
for gap_ms in GAPS_MS:
    a.gap_ms = gap_ms
    gap = int(a.gap_ms * 22050 / 1000)
    mask = torch.ones(1, L, device=dev)
    mask[..., L // 2 - gap // 2: L // 2 - gap // 2 + gap] = 0      # tester_inpainting.py:231-242

    def run():
        if a.config == "inpaint":
            return smp.predict_inpainting(y * mask, mask)
        return smp.predict_unconditional((B, L), dev)

    net._ensure_weights(dev)
    smp.calls = 0                     # every gap of a sweep starts from the same noise key
    smp.sampler.nb_steps = 2          # warm-up: module loading, workspace, graph capture (the graphs do not depend on the step count)
    run()
    smp.sampler.nb_steps = a.steps
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = run()                       # includes the result gather (the only collective)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt_wall = time.perf_counter() - t0
    tt = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt)                    # device time, max over ranks
    evals = 2 * a.steps - 1
    ok = bool(torch.isfinite(out).all()) and tuple(out.shape) == (B, L)
    if a.config == "inpaint":
        keep = mask[0].bool().clone()
        keep[L // 2 - gap // 2 - 100: L // 2 + gap // 2 + 100] = False
        ok = ok and bool(torch.allclose(out[:, keep], (y * mask)[:, keep], atol=1e-5))
    if rank == 0:
        lo, hi = shard_bounds(B, 0, world)
        print(json.dumps({"config": a.config, "batch": B, "len": L, "gap_ms": a.gap_ms if a.config == "inpaint" else None, "n_gpus": world,
                          "sampler_steps": a.steps, "denoiser_evals": evals, "seconds": dt, "seconds_wall": dt_wall, "clips_per_s": B / dt,
                          "seconds_per_eval": dt / evals, "clips_per_rank": hi - lo, "output_ok": ok, "conv_mode": a.conv_mode,
                          "noise": a.noise, "cuda_graph": a.noise == "device" and not a.no_graph, "xi": a.xi,
                          "timing": "CUDA events around the whole predict call incl. the result gather, max over ranks; 1 warm-up call of 2 steps before"}))
if world > 1:
    dist.destroy_process_group()
