#!/bin/bash
# evidence for the last two fused kernels + the sampler configs with the final build
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"out_block_kernel|gn_act_tc2_kernel<false, true>|gn_act_tc2_kernelILb0ELb1" -c 6 -f -o gpurun_out/r2_ncu_out_up_full python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_f4.log 2>&1
echo "ncu full rc=$?"
timeout 300 python tools/bench_sampler.py --config inpaint --batch 32 --gap-ms 300 --steps 35 > gpurun_out/r2_sampler_config3_inpaint_b32_v4.json 2> gpurun_out/r2_samp.err; cat gpurun_out/r2_sampler_config3_inpaint_b32_v4.json
timeout 300 python tools/bench_sampler.py --config uncond --batch 8 --steps 35 > gpurun_out/r2_sampler_config2_uncond_b8_v4.json 2>> gpurun_out/r2_samp.err; cat gpurun_out/r2_sampler_config2_uncond_b8_v4.json
