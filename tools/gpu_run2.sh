#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 13 5 9 15; do
  AID_TC_DEBUG=$dbg TC_SHAPES="8,64,64,4096,2;8,96,192,1024,4;8,128,256,512,16;8,256,384,128,64" python tools/time_conv.py 3 5x3 2>&1 | grep TFLOP
done > gpurun_out/r2_ablate.log 2>&1
cat gpurun_out/r2_ablate.log
python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2_t2_all.log 2>&1
echo "all tests rc=$?"
tail -25 gpurun_out/r2_t2_all.log
