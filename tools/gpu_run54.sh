#!/bin/bash
# final numbers of the round: bench line, sampler configs 2 / 3, guided clip (taped path, un-fused blocks)
mkdir -p gpurun_out
timeout 600 python bench.py 2>gpurun_out/r2_bench_v3.err | tail -1 | tee gpurun_out/r2_bench_tc2_v3.json
timeout 300 python tools/bench_sampler.py --config inpaint --batch 32 --gap-ms 300 --steps 35 > gpurun_out/r2_sampler_config3_inpaint_b32_v3.json 2> gpurun_out/r2_samp.err; cat gpurun_out/r2_sampler_config3_inpaint_b32_v3.json
timeout 300 python tools/bench_sampler.py --config uncond --batch 8 --steps 35 > gpurun_out/r2_sampler_config2_uncond_b8_v3.json 2>> gpurun_out/r2_samp.err; cat gpurun_out/r2_sampler_config2_uncond_b8_v3.json
timeout 300 python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 35 --xi 0.25 > gpurun_out/r2_sampler_guided_xi025_b1_v3.json 2>> gpurun_out/r2_samp.err; cat gpurun_out/r2_sampler_guided_xi025_b1_v3.json
tail -3 gpurun_out/r2_samp.err
