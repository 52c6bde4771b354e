#!/bin/bash
mkdir -p gpurun_out
SH="8,64,64,4096,2;8,96,128,2048,4;8,128,256,512,16;8,256,384,128,64"
for dbg in 2048 2049 2061 2063; do
  AID_TC_DEBUG=$dbg TC_SHAPES="$SH" python tools/time_conv.py 3 5x3 2>&1 | grep -v Warn
done > gpurun_out/r2_prof.log 2>&1
for na in 2 4; do echo "NA=$na"; AID_TC2_NA=$na AID_TC_DEBUG=2048 TC_SHAPES="$SH" python tools/time_conv.py 3 5x3 2>&1 | grep -v Warn; done >> gpurun_out/r2_prof.log 2>&1
cat gpurun_out/r2_prof.log
