#!/bin/bash
mkdir -p gpurun_out
python tools/bench_vjp.py --len 262144 > gpurun_out/r2_vjp_22k.json 2> gpurun_out/r2_vjp.err
python tools/bench_vjp.py --net paper_44k --len 184184 > gpurun_out/r2_vjp_44k.json 2>> gpurun_out/r2_vjp.err
cat gpurun_out/r2_vjp_22k.json gpurun_out/r2_vjp_44k.json; tail -3 gpurun_out/r2_vjp.err
python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 35 --xi 0.25 > gpurun_out/r2_samp_guided_b1.json 2>> gpurun_out/r2_vjp.err
cat gpurun_out/r2_samp_guided_b1.json
python tools/bench_sampler.py --config inpaint --batch 32 --gap-ms 300 --steps 35 > gpurun_out/r2_sampler_config3_inpaint_b32.json 2>> gpurun_out/r2_vjp.err
cat gpurun_out/r2_sampler_config3_inpaint_b32.json
python tools/bench_sampler.py --config uncond --batch 8 --steps 35 > gpurun_out/r2_sampler_config2_uncond_b8.json 2>> gpurun_out/r2_vjp.err
cat gpurun_out/r2_sampler_config2_uncond_b8.json; tail -3 gpurun_out/r2_vjp.err
