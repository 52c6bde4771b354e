#!/bin/bash
# bench lines of the round-2 build on one GPU (default, and the reference arm as the driver runs it)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_tc2.json 2> gpurun_out/r2_bench_tc2.err; tail -c 3200 gpurun_out/r2_bench_tc2.json | head -c 2400
