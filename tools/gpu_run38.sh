#!/bin/bash
# final round-2 build: launch list at B = 8, unconditional sampling (config 2), guided sampling, VJP timings
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2d_launches_tc2_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2d_ncu_ll.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2d_launches_tc2_b8.csv "one forward, B=8 x 262144, conv_mode 2 (round 2 final: fused 64 / 96-channel layers, static epilogue, cta_group::2 on the 256-cout layers)" > gpurun_out/r2d_launches_tc2_b8.summary.txt; head -12 gpurun_out/r2d_launches_tc2_b8.summary.txt
timeout 300 python tools/bench_sampler.py --config uncond --batch 8 --steps 35 > gpurun_out/r2d_sampler_config2.json 2>/dev/null; cat gpurun_out/r2d_sampler_config2.json
timeout 300 python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 35 --xi 0.25 > gpurun_out/r2d_sampler_guided.json 2>/dev/null; cat gpurun_out/r2d_sampler_guided.json
timeout 300 python tools/bench_vjp.py > gpurun_out/r2d_vjp_22k.json 2>/dev/null; cat gpurun_out/r2d_vjp_22k.json
timeout 300 python tools/bench_vjp.py --net paper_44k --len 184184 > gpurun_out/r2d_vjp_44k.json 2>/dev/null; cat gpurun_out/r2d_vjp_44k.json
