#!/bin/bash
mkdir -p gpurun_out
for dbg in 2048 2063; do
  AID_TC_DEBUG=$dbg TC_SHAPES="8,64,64,4096,2;8,128,256,512,16;8,256,384,128,64" python tools/time_conv.py 3 5x3 2>&1 | grep -v Warn
done > gpurun_out/r2_prof3.log 2>&1
cat gpurun_out/r2_prof3.log
