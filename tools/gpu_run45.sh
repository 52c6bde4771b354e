#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 35 --xi 0.25 > gpurun_out/r2e_sampler_guided.json 2>/dev/null; cat gpurun_out/r2e_sampler_guided.json
timeout 300 python tools/bench_vjp.py > gpurun_out/r2e_vjp_22k.json 2>/dev/null
timeout 300 python tools/bench_vjp.py --net paper_44k --len 184184 > gpurun_out/r2e_vjp_44k.json 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_launches_vjp_b1.csv python tools/profile_vjp.py --batch 1 > gpurun_out/r2e_ncu_vjp.log 2>&1
python tools/summarize_launches.py gpurun_out/r2e_launches_vjp_b1.csv "backward (VJP) of one clip x 262144, conv_mode 2, 1x1 data gradients on tcgen05" > gpurun_out/r2e_launches_vjp_b1.summary.txt; head -14 gpurun_out/r2e_launches_vjp_b1.summary.txt
