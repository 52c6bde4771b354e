#!/bin/bash
# fused out block: block-level parity + timing, bench-shape invariants + golden, forward time with / without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_out_block.py -m gpu -q -s 2>&1 | grep "fused\|passed\|failed\|Error\|assert" | tail -24
timeout 900 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_unet.py -m gpu -x -q -s 2>&1 | grep "whole batch\|rel-L2\|passed\|failed\|Error\|assert" | head -30
echo "== fused"; timeout 300 python tools/time_forward.py 1 8 32
echo "== AID_OUT_FUSED=0"; AID_OUT_FUSED=0 timeout 300 python tools/time_forward.py 1 8 32
