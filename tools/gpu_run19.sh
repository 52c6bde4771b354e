#!/bin/bash
# cta_group::2 bring-up: parity tests first (bounded), then per-layer timings with the pair mode on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --timeout 90 --timeout-method thread -k "cta_pair or single_cta" 2>&1 | tail -15
echo "== cg2 on"
AID_TC2_CG2=1 AID_TC_DEBUG=4096 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -14
echo "== cg2 off"
AID_TC2_CG2=0 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -8
echo "== cg2 on, nA 4"
AID_TC2_CG2=1 AID_TC2_NA=4 timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -8
echo "== profile cg2 on"
AID_TC2_CG2=1 AID_TC_DEBUG=2048 TC_SHAPES="8,64,64,4096,2;8,128,256,512,16;8,256,384,128,64" timeout 120 python tools/time_conv.py 3 2>&1 | tail -8
echo "== profile cg2 off"
AID_TC2_CG2=0 AID_TC_DEBUG=2048 TC_SHAPES="8,64,64,4096,2;8,128,256,512,16;8,256,384,128,64" timeout 120 python tools/time_conv.py 3 2>&1 | tail -8
