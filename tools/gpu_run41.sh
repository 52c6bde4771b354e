#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_tc2.json 2> gpurun_out/r2_bench_tc2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_tc2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['fused_layers']['frac'], d['fused_layers']['kernel_ms_per_step'], d['roofline']['dilated_layers_all'], d['clocks'])"
