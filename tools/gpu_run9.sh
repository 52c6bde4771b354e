#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vjp.py -m gpu -q -s > gpurun_out/r2_t9.log 2>&1
echo "vjp tests rc=$?"; grep -v "^$" gpurun_out/r2_t9.log | tail -60
