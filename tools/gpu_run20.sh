#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --timeout 90 --timeout-method thread -k "cta_pair or single_cta" 2>&1 | tail -5
echo "== cg2 on"
AID_TC2_CG2=1 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -8
echo "== cg2 off"
AID_TC2_CG2=0 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -8
echo "== cg2 on, epilogue body off"
AID_TC2_CG2=1 AID_TC_DEBUG=1 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -8
echo "== cg2 off, epilogue body off"
AID_TC2_CG2=0 AID_TC_DEBUG=1 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -8
echo "== profile cg2 on"
AID_TC2_CG2=1 AID_TC_DEBUG=2048 TC_SHAPES="8,64,64,4096,2;8,96,128,2048,4;8,128,256,512,16;8,256,384,128,64" timeout 120 python tools/time_conv.py 3 2>&1 | tail -8
