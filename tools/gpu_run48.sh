#!/bin/bash
# fused init block (conv_init.cu): parity at the bench shape and the oracle tests, then the bench line with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_unet.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
echo "== bench, fused init blocks"
timeout 600 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bench_initfused.json
echo "== bench, AID_INIT_FUSED=0"
AID_INIT_FUSED=0 timeout 600 python bench.py --steps 8 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bench_initunfused.json
