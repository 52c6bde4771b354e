#!/bin/bash
# BASELINE config 5: gap sweep 25-1500 ms, 35-step EDM, batch 64 over 4 GPUs; plus the 2-rank NCCL test of ShardedSampler
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/bench_sampler.py --config inpaint --batch 64 --gap-sweep --steps 35 > gpurun_out/r2_sampler_config5_sweep_b64_n4.jsonl 2> gpurun_out/r2_n4.err
echo "config5 rc=$?"; cat gpurun_out/r2_sampler_config5_sweep_b64_n4.jsonl; tail -3 gpurun_out/r2_n4.err
python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r2_t_multi.log 2>&1
echo "multi test rc=$?"; tail -5 gpurun_out/r2_t_multi.log
