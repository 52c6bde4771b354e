#!/usr/bin/env python
"""Benchmark of the denoiser forward (BASELINE.json metric): clips/s at 22.05 kHz, 262144 samples, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one forward of the 186.3 M-parameter CQT-octave U-Net over one batch of 32 synthetic clips per GPU
(random-init weights, sigma shared by the batch, the way the sampler calls it).  One JSON line on rank 0:
  value       clips/s with the batch resident in HBM, CUDA events, max over ranks, whole job
  e2e         the same through the public nn.Module call with pinned HOST buffers (H2D + forward + D2H per step)
  roofline    dominant kernel (the dilated 5x3 convolutions): algorithmic FLOPs / CUDA-event time vs measured peak
  cpu_baseline the reference's CPU path on a bounded sample (one clip of the full length), rank 0, N=1 only
  fp32_grade  the same step in conv_mode 1 (split fp16 x3, <= 1e-4 from the fp32 reference), N=1 only
`--impl reference` times the reference's CPU implementation of the path on the host cores, on the same workload definition:
the reference's own unet.py when build() could copy it into the git-ignored oracle/_ref/ (kind "reference"; the un-vendored
CQT dependency comes from oracle/cqt_oracle.py), else the oracle restatement (kind "port").  That arm never imports the
product package and never loads libaid_b200.so.
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


class _Args(dict):
    """Attribute dict for the reference's `args.network.*` / `args.exp.*` reads (unet.py:595-655); no product import."""
    __getattr__ = dict.__getitem__

    @staticmethod
    def wrap(o):
        return _Args({k: _Args.wrap(v) for k, v in o.items()}) if isinstance(o, dict) else o


def paper_args(L):
    """conf/network/paper_1912_unet_cqt_oct_attention_adaLN_2.yaml + conf/exp/maestro22k_*.yaml as the reference reads them."""
    return _Args.wrap({
        "exp": {"sample_rate": 22050, "audio_len": L},
        "network": {"use_fencoding": False, "use_norm": True, "emb_dim": 256, "Ns": [64, 96, 96, 128, 128, 256, 256], "Ss": [2] * 6,
                    "num_dils": [2, 3, 4, 5, 6, 7, 7], "attention_layers": [0, 0, 0, 0, 1, 1, 1, 1], "bottleneck_type": "res_dil_convs",
                    "num_bottleneck_layers": 1, "cqt": {"window": "kaiser", "beta": 1, "num_octs": 7, "bins_per_oct": 64},
                    "attention_dict": {"num_heads": 8, "attn_dropout": 0.0, "bias_qkv": False, "N": 0, "rel_pos_num_buckets": 32,
                                       "rel_pos_max_distance": 64, "use_rel_pos": False, "Nproj": 8}},
    })


def cpu_denoiser(L):
    """The reference's CPU implementation of the path -> (forward(x[B,L], c_noise[1,1]), kind, description).

    kind "reference": the reference's own unet.py (build() copies it, unmodified, into the git-ignored oracle/_ref/ when
    /root/reference exists; its un-vendored CQT dependency is supplied by oracle/cqt_oracle.py), default-initialised by its own
    constructor.  kind "port": oracle/unet_oracle.py (the restatement; its resampler is a depthwise conv1d instead of the
    reference's dense [F,F,8] weight, so it is somewhat FASTER than the real reference) with random weights of the schema
    stored in tests/golden/golden_meta.json.  Neither imports the product package or loads libaid_b200.so."""
    import torch
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_dir, "networks", "unet_cqt_oct_with_projattention_adaLN_2.py")):
        import cqt_oracle
        cqt_oracle.install_as_cqt_nsgt_pytorch()
        sys.path.insert(0, ref_dir)
        from networks.unet_cqt_oct_with_projattention_adaLN_2 import Unet_CQT_oct_with_attention as RefNet
        torch.manual_seed(1234)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):     # the constructor prints its layer plan
            net = RefNet(paper_args(L), "cpu")

        def fwd(x, cn):
            with torch.no_grad():
                return net(x, cn)
        return fwd, "reference", "the reference's own unet.py (oracle/_ref copy) + the CQT restatement, default init"
    import unet_oracle
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_meta.json")))
    g = torch.Generator().manual_seed(1234)
    sd = {}
    for name, shape in meta["schema_paper"]:
        fan = 1
        for d in shape[1:]:
            fan *= d
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / max(fan, 1) ** 0.5
    cfg = dict(num_octs=7, bins_per_oct=64, sample_rate=22050, audio_len=L, window="kaiser", beta=1, Ns=[64, 96, 96, 128, 128, 256, 256],
               num_dils=[2, 3, 4, 5, 6, 7, 7], attention_layers=[0, 0, 0, 0, 1, 1, 1, 1])
    orc = unet_oracle.UnetOracle(cfg, sd)
    return (lambda x, cn: orc(x, cn)), "port", "oracle/unet_oracle.py (restatement of the reference's PyTorch path), random weights"


def bench_config(B, L):
    """The workload both arms report (the reference arm echoes it verbatim)."""
    return {"workload": f"denoiser forward, paper_1912 CQT-octave U-Net (186.3M params, random init), batch {B} x {L} samples per GPU, shared sigma",
            "batch_per_gpu": B, "audio_len": L,
            "l2": "per-step working set is tens of GB of activations, far larger than the 126 MB L2 (no flush needed)"}


METRIC = "denoiser-fwd clips/s at 22.05 kHz, 262144-len, batch 32"


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on all host cores.  A step is a bounded sample of
    the batch-B workload (n clips of the full length per forward); value = clips/s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    L, B = args.len, args.batch
    fwd, kind, what = cpu_denoiser(L)
    cn = torch.tensor([[-0.3]])
    g = torch.Generator().manual_seed(0)
    x1 = torch.randn(1, L, generator=g) * 0.5
    t0 = time.perf_counter(); fwd(x1, cn); t_first = time.perf_counter() - t0          # also pages everything in
    # clips per step: 2 when the whole run (W + K steps) stays within ~5 minutes, else 1; K is cut only if even that is too long
    total = args.steps + args.warmup
    n = 2 if 2 * t_first * total <= 300.0 else 1
    n = min(n, B)
    k_eff = args.steps if n * t_first * total <= 600.0 else max(1, int(600.0 / (n * t_first)) - args.warmup)
    x = torch.randn(n, L, generator=g) * 0.5
    for _ in range(args.warmup):
        fwd(x, cn)
    ts = []
    for _ in range(k_eff):
        t0 = time.perf_counter(); fwd(x, cn); ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    val = n / (ms / 1e3)
    sample = (f"{what}; each step = one forward over {n} clip(s) x {L} samples (a bounded sample of the batch-{B} step: "
              f"a full step would be {B / n:.0f}x longer), {k_eff} timed steps after {args.warmup} warm-up steps, {threads} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": k_eff, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(B, L), "sample_clips_per_step": n,
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": threads, "kind": kind, "sample": sample},
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
    ap.add_argument("--no-fp32-grade", action="store_true", help="skip the secondary conv_mode 1 measurement")
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
    n_f, t_f, fl_f, by_f = C.c_uint64(), C.c_double(), C.c_double(), C.c_double()      # fused dilated layers (conv_comb.cu)
    _lib.check(lib.aid_profile_read(net._handle, 2, C.byref(n_f), C.byref(t_f), C.byref(fl_f), C.byref(by_f)), net._handle)
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

    fp32_grade = None
    if world == 1 and args.conv_mode == 2 and not args.no_fp32_grade:
        # the same step in conv_mode 1 (error-compensated split fp16, 3 MMAs per tap: <= 1e-4 from the fp32 reference)
        net._release()
        torch.cuda.empty_cache()
        cfg1 = aid_b200.paper_22k(L, conv_mode=1)
        net1 = aid_b200.Unet_CQT_oct_with_attention(cfg1, dev)
        net1.load_state_dict(aid_b200.random_state_dict(cfg1, seed=1234))
        for _ in range(3):
            net1.denoise_fused(x, cn, out=out)
        k1 = min(args.steps, 5)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(k1):
            net1.denoise_fused(x, cn, out=out)
        g1.record()
        torch.cuda.synchronize()
        ms1 = g0.elapsed_time(g1) / k1
        fp32_grade = {"conv_mode": 1, "dtype": "f16x3-split (fp32 accumulate), <= 1e-4 rel-L2 from the fp32 reference", "value": B / (ms1 / 1e3),
                      "unit": "clips/s", "ms_per_step": ms1, "steps": k1, "warmup": 3}
        net1._release()
        del net1

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
            "metric": METRIC, "value": world * B / (ms / 1e3), "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {0: "f32", 1: "f16x3-split (fp32 accumulate)", 2: "f16 (fp32 accumulate)"}[args.conv_mode],
            "data": "synthetic",
            "config": bench_config(B, L), "conv_mode": args.conv_mode,
            "parallelism": f"batch sharded over {world} rank(s), {B} clips each, no collective inside the step",
            "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": B * L * 4, "d2h_bytes_per_step": B * L * 4},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": {0: "dilated 5x3 conv residual layers (conv_simt_kernel<5,3,8>)", 1: "dilated 5x3 conv (tcgen05, conv_tc_kernel, 3 MMAs per tap)",
                                    2: "dilated 5x3 conv of the 128 / 256-channel residual layers (tcgen05, 1 fp16 MMA per tap: conv_tc2_kernel, conv_tc2_cg2_kernel); FLOPs = executed taps only"}[args.conv_mode],
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
        if t_f.value > 0:
            # the 64 / 96-channel residual layers run as ONE kernel each (normalise + modulate + GELU + convert + convolve + epilogue):
            # their bound is the layer's HBM traffic (4 B read + 4 B written per element), reported next to their tensor throughput
            gbs = (by_f.value / 1e9) / (t_f.value / 1e3)
            line["fused_layers"] = {"kernel": "conv_comb_kernel / conv_comb96_kernel (fused dilated residual layers, 64 / 96 channels)",
                                    "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                                    "tflops": (fl_f.value / 1e12) / (t_f.value / 1e3), "launches_timed": int(n_f.value),
                                    "traffic": (json.load(open(prof_json)).get("fused_layers_mode2_bytes_per_launch") if os.path.exists(prof_json) else None),
                                    "kernel_ms_per_step": t_f.value / args.steps, "kernel_share_of_step": (t_f.value / args.steps) / ms}
            line["roofline"]["dilated_layers_all"] = {"tflops": ((fl.value + fl_f.value) / 1e12) / ((t_ms.value + t_f.value) / 1e3),
                                                      "kernel_ms_per_step": (t_ms.value + t_f.value) / args.steps,
                                                      "kernel_share_of_step": ((t_ms.value + t_f.value) / args.steps) / ms}
        if fp32_grade is not None:
            line["fp32_grade"] = fp32_grade
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload on the host cores: one clip of the full length, 1 warm-up + 1 timed forward
            del net
            torch.cuda.empty_cache()
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            fwd, kind, what = cpu_denoiser(L)
            xs = torch.randn(1, L, generator=torch.Generator().manual_seed(0)) * 0.5
            cns = torch.tensor([[-0.3]])
            fwd(xs, cns)
            t0 = time.perf_counter(); fwd(xs, cns); tt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / tt, "unit": "clips/s", "cores": threads, "kind": kind,
                                    "sample": f"{what}; 1 clip x {L} samples (one of the {B} clips of a step), 1 warm-up + 1 timed forward, {threads} threads"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
