#!/bin/bash
# bench lines of the round-2 build on one GPU (default, and the reference arm as the driver runs it)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_tc2.json 2> gpurun_out/r2_bench_tc2.err; tail -c 2500 gpurun_out/r2_bench_tc2.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 800 gpurun_out/r2_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
