#!/bin/bash
# two ranks on one box: the NCCL test of ShardedSampler and the forward bench as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['n_gpus'], d['e2e']['value'], d['clocks'])"
