#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --timeout 90 --timeout-method thread -k "single_fp16 or cta_pair or single_cta or 1x1" 2>&1 | tail -5
echo "== fast epilogue"
timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -6
echo "== generic epilogue"
AID_TC2_FASTEPI=0 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -6
echo "== fast epilogue + cg2"
AID_TC2_CG2=1 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -6
