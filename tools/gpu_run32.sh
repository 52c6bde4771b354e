#!/bin/bash
# final state of round 2: full GPU suite, launch list at B = 8, ncu full captures of the fused kernels, sampler config 3, guided sampling
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread 2>&1 | tail -4
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches_tc2_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2c_ncu_ll.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2c_launches_tc2_b8.csv "one forward, B=8 x 262144, conv_mode 2 (round 2 final: fused 64 / 96-channel layers, static epilogue, cta_group::2 on the 256-cout layers)" > gpurun_out/r2c_launches_tc2_b8.summary.txt; cat gpurun_out/r2c_launches_tc2_b8.summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_comb96_kernel -s 2 -c 2 -f -o gpurun_out/r2_ncu_conv_comb96_full python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_f3.log 2>&1
echo "ncu comb96 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2_cg2_kernel -s 10 -c 2 -f -o gpurun_out/r2_ncu_conv_tc2_cg2_full python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_f4.log 2>&1
echo "ncu cg2 rc=$?"
timeout 600 python tools/bench_sampler.py --config inpaint --batch 32 --gap-ms 300 --steps 35 > gpurun_out/r2c_sampler_config3_inpaint_b32.json 2> gpurun_out/r2c_s3.err; cat gpurun_out/r2c_sampler_config3_inpaint_b32.json
