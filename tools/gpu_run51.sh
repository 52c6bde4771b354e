#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_init_block.py -m gpu -q -s 2>&1 | grep "fused\|passed\|failed" | tail -16
timeout 300 python tools/time_forward.py 1 8
AID_INIT_FUSED=0 timeout 300 python tools/time_forward.py 1 8
