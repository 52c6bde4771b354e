#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_shape.py tests/test_gpu_unet.py -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_fastepi.json 2> gpurun_out/r2_bench_fastepi.err; tail -c 1500 gpurun_out/r2_bench_fastepi.json
