#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -s -k attention > gpurun_out/r2_t15.log 2>&1
echo "attention tests rc=$?"; grep "tcgen05 attention\|passed\|failed\|^E  " gpurun_out/r2_t15.log | head -30
timeout 300 python tools/time_attention.py 2>&1 | tail -10
