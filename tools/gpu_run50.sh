#!/bin/bash
# thin-in convolutions with 4 output channels per pass: parity tests, forward time with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py -m gpu -x -q 2>&1 | tail -3
echo "== CO4"; timeout 300 python tools/time_forward.py 1 8 32
echo "== CO1"; AID_THIN_CO4=0 timeout 300 python tools/time_forward.py 1 8 32
