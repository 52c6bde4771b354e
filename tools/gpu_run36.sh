#!/bin/bash
# BASELINE config 4 with the final round-2 build: inpainting, 1500 ms gap, batch 256 sharded over 8 GPUs, NCCL gather inside the timed region
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_sampler.py --config inpaint --batch 256 --gap-ms 1500 --steps 35 > gpurun_out/r2c_sampler_config4_b256_n8.json 2> gpurun_out/r2c_n8.err
echo "config4 rc=$?"; cat gpurun_out/r2c_sampler_config4_b256_n8.json; tail -2 gpurun_out/r2c_n8.err
