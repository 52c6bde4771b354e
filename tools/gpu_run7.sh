#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc2_layout.py -m gpu -q -x > gpurun_out/r2_t7.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r2_t7.log
for dbg in 2048 0 1; do
  AID_TC_DEBUG=$dbg python tools/time_conv.py 3 5x3 2>&1 | grep -v Warn
done > gpurun_out/r2_prof4.log 2>&1
cat gpurun_out/r2_prof4.log
