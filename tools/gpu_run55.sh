#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_unet.py -m gpu -x -q -s 2>&1 | grep "whole batch\|row \|passed\|failed\|Error\|assert" | head -30
echo "== fused upsampling"; timeout 300 python tools/time_forward.py 1 8 32
echo "== AID_UP_FUSED=0"; AID_UP_FUSED=0 timeout 300 python tools/time_forward.py 1 8 32
