#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 13; do
  AID_TC_DEBUG=$dbg python tools/time_conv.py 3 5x3 2>&1 | grep TFLOP
done > gpurun_out/r2_ablate2.log 2>&1
cat gpurun_out/r2_ablate2.log
python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc2_layout.py tests/test_gpu_bench_shape.py -m gpu -q -x > gpurun_out/r2_t3.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_t3.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32-grade > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
python -c "
import json; d = json.load(open('gpurun_out/r2_bench2.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks'])"
