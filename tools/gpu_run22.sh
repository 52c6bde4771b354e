#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --timeout 90 --timeout-method thread -k "single_fp16 or cta_pair or single_cta or 1x1" 2>&1 | tail -5
for ew in 16 8; do
echo "== EW $ew"
AID_TC2_EW=$ew timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -6
done
echo "== EW 16 cg2"
AID_TC2_CG2=1 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -6
echo "== EW 16 1x1"
timeout 120 python tools/time_conv.py 3 1x1 2>&1 | tail -7
echo "== EW 8 1x1"
AID_TC2_EW=8 timeout 120 python tools/time_conv.py 3 1x1 2>&1 | tail -7
