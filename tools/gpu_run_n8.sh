#!/bin/bash
# BASELINE config 4: inpainting, 1500 ms gap, batch 256 sharded over 8 GPUs, one NCCL gather at the end (inside the timed region)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_sampler.py --config inpaint --batch 256 --gap-ms 1500 --steps 35 > gpurun_out/r2_sampler_config4_b256_n8.json 2> gpurun_out/r2_n8.err
echo "config4 rc=$?"; cat gpurun_out/r2_sampler_config4_b256_n8.json; tail -3 gpurun_out/r2_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_sampler.py --config inpaint --batch 256 --gap-ms 1500 --steps 35 --noise host > gpurun_out/r2_sampler_config4_b256_n8_hostnoise.json 2>> gpurun_out/r2_n8.err
cat gpurun_out/r2_sampler_config4_b256_n8_hostnoise.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2>> gpurun_out/r2_n8.err
tail -c 600 gpurun_out/r2_bench_n8.json
