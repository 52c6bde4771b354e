#!/usr/bin/env python
"""Benchmark of the denoiser forward (BASELINE.json metric): clips/s at 22.05 kHz, 262144 samples, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one forward of the 186.3 M-parameter CQT-octave U-Net over one batch of 32 synthetic clips per GPU
(random-init weights, sigma shared by the batch, the way the sampler calls it).  One JSON line on rank 0:
  value       clips/s with the batch resident in HBM, CUDA events, max over ranks, whole job
  e2e         the same through the public nn.Module call with pinned HOST buffers (H2D + forward + D2H per step)
  roofline    dominant kernel (the dilated 5x3 convolutions): algorithmic FLOPs / CUDA-event time vs measured peak
  cpu_baseline the CPU oracle (port of the reference's PyTorch path) on a bounded sample, rank 0, N=1 only
`--impl reference` times the reference's CPU path (the oracle port: the reference is Python + an un-vendored
dependency and cannot travel to the GPU box) on the host cores, on the same workload definition.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

GFLOP_PER_CLIP = {262144: 4072.0, 65536: 1018.0}   # SURVEY.md section 8d (algorithmic, whole forward)
GB_PER_CLIP = {262144: 14.5, 65536: 3.62}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], tensor_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  The process is started before the warm-up (its start-up takes
    a few hundred ms); only the samples that arrive between mark_start() and mark_end() -- the timed region -- are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.t0, self.t1 = [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in list(self.rows):
            if (self.t0 is not None and ts < self.t0) or (self.t1 is not None and ts > self.t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_forward_time(L, n_clips, threads):
    """Seconds per forward of the CPU oracle (the reference's PyTorch path restated) on n_clips x L, best of 2."""
    import torch
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import aid_b200
    from util import make_oracle
    torch.set_num_threads(threads)
    cfg = aid_b200.paper_22k(L)
    orc = make_oracle(cfg, aid_b200.random_state_dict(cfg, seed=1234))
    x = torch.randn(n_clips, L, generator=torch.Generator().manual_seed(0)) * 0.5
    cn = torch.tensor([[-0.3]])
    return orc, x, cn


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    L, B = args.len, args.batch
    orc, x, cn = oracle_forward_time(L, 1, threads)
    # a step = one clip of the batch-32 workload (bounded sample); no JIT or clocks to warm on the CPU, so one warm-up
    t0 = time.perf_counter(); orc(x, cn); t_first = time.perf_counter() - t0
    k_eff = max(1, min(args.steps, int(240.0 / max(t_first, 1e-3))))
    ts = []
    for _ in range(k_eff):
        t0 = time.perf_counter(); orc(x, cn); ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    val = 1.0 / (ms / 1e3)
    sample = f"1 clip x {L} samples per step (of the batch-{B} workload), {k_eff} timed steps after 1 warm-up"
    print(json.dumps({
        "impl": "reference", "metric": "denoiser-fwd clips/s at 22.05 kHz, 262144-len, batch 32", "value": val, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": k_eff, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"denoiser forward, paper_1912 CQT-octave U-Net (186.3M params), batch {B} x {L} samples per GPU",
                   "batch_per_gpu": B, "audio_len": L},
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--len", type=int, default=262144)
    ap.add_argument("--conv-mode", type=int, default=int(os.environ.get("AID_CONV_MODE", "2")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import aid_b200
    from aid_b200 import _lib

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L = args.batch, args.len

    cfg = aid_b200.paper_22k(L, conv_mode=args.conv_mode)
    net = aid_b200.Unet_CQT_oct_with_attention(cfg, dev)
    net.load_state_dict(aid_b200.random_state_dict(cfg, seed=1234))
    net._ensure_weights(dev)
    lib = _lib.lib()
    g = torch.Generator().manual_seed(100 + rank)
    x_host = (torch.randn(B, L, generator=g) * 0.5).pin_memory()
    out_host = torch.empty(B, L).pin_memory()
    x = x_host.to(dev)
    cn = torch.tensor([[-0.3]], device=dev)
    out = torch.empty_like(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        net.denoise_fused(x, cn, out=out)

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    lib.aid_profile(net._handle, 1)
    launches0 = lib.aid_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if clocks:
        clocks.mark_start()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    if clocks:
        clocks.mark_end()
    ms = e0.elapsed_time(e1) / args.steps
    launches = lib.aid_launch_count() - launches0
    n_l, t_ms, fl, by = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()
    _lib.check(lib.aid_profile_read(net._handle, 0, C.byref(n_l), C.byref(t_ms), C.byref(fl), C.byref(by)), net._handle)
    lib.aid_profile(net._handle, 0)
    clk = clocks.stop() if clocks else None

    # end to end through the public API: pinned host input -> device -> forward -> pinned host output, every step
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        y = net(xd, cn)
        out_host.copy_(y, non_blocking=True)

    e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1) / args.steps

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank == 0:
        pk = peaks()
        conv_tflops = (fl.value / 1e12) / (t_ms.value / 1e3) if t_ms.value > 0 else 0.0
        prof_json = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = None
        if os.path.exists(prof_json):
            try:
                traffic = json.load(open(prof_json)).get(f"conv5x3_mode{args.conv_mode}_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "denoiser-fwd clips/s at 22.05 kHz, 262144-len, batch 32", "value": world * B / (ms / 1e3), "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {0: "f32", 1: "f16x3-split (fp32 accumulate)", 2: "f16 (fp32 accumulate)"}[args.conv_mode],
            "data": "synthetic",
            "config": {"workload": f"denoiser forward, paper_1912 CQT-octave U-Net (186.3M params, random init), batch {B} x {L} samples per GPU, shared sigma",
                       "batch_per_gpu": B, "audio_len": L, "conv_mode": args.conv_mode,
                       "parallelism": f"batch sharded over {world} rank(s), no collective inside the step",
                       "l2": "per-step working set is tens of GB of activations, far larger than the 126 MB L2 (no flush needed)"},
            "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * L * 4},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": {0: "dilated 5x3 conv residual layers (conv_simt_kernel<5,3,8>)", 1: "dilated 5x3 conv (tcgen05, conv_tc_kernel, 3 MMAs per tap)",
                                    2: "dilated 5x3 conv (tcgen05, conv_tc2_kernel, 1 fp16 MMA per tap)"}[args.conv_mode],
                         "achieved": conv_tflops, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": conv_tflops / pk["tensor"],
                         "peak_source": f"{pk['src']} bf16 dense GEMM, sustained", "traffic": traffic,
                         "algorithmic_gbs": (by.value / 1e9) / (t_ms.value / 1e3) if t_ms.value > 0 else 0.0,
                         "frac_of_hbm_peak": ((by.value / 1e9) / (t_ms.value / 1e3) / pk["hbm"]) if t_ms.value > 0 else 0.0,
                         "launches_timed": int(n_l.value), "kernel_ms_per_step": t_ms.value / args.steps,
                         "kernel_share_of_step": (t_ms.value / args.steps) / ms,
                         "whole_forward": {"tflops": B * GFLOP_PER_CLIP.get(L, 0) / 1e3 / (ms / 1e3),
                                           "frac_of_tensor_roof": B * GFLOP_PER_CLIP.get(L, 0) / 1e3 / (ms / 1e3) / pk["tensor"],
                                           "algorithmic_gbs": B * GB_PER_CLIP.get(L, 0) / (ms / 1e3),
                                           "frac_of_hbm_roof": B * GB_PER_CLIP.get(L, 0) / (ms / 1e3) / pk["hbm"]}},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            Ls = 65536  # bounded sample: one clip of a quarter of the length (work per sample is length-independent: 15.5 MFLOP)
            orc, xs, cns = oracle_forward_time(Ls, 1, threads)
            orc(xs, cns)
            t0 = time.perf_counter(); orc(xs, cns); tt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": (Ls / L) / tt, "unit": "clips/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle (reference PyTorch path restated) on 1 clip x {Ls} samples, 1 warm-up + 1 timed forward, "
                                              f"scaled by {Ls}/{L} to 262144-sample clips (work per sample is constant)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
