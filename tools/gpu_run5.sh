#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc2_layout.py -m gpu -q -x > gpurun_out/r2_t5.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_t5.log
for dbg in 2048 0 1 13; do
  AID_TC_DEBUG=$dbg python tools/time_conv.py 3 5x3 2>&1 | grep -v Warn
done > gpurun_out/r2_prof2.log 2>&1
cat gpurun_out/r2_prof2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32-grade > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python -c "
import json; d = json.load(open('gpurun_out/r2_bench3.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks'])"
