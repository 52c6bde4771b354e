#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_vjp.py 2>&1 | grep -v "Attention layer\|^$" | tail -12
timeout 900 python -m pytest tests/test_gpu_vjp.py -m gpu -q -s -k "denoiser_vjp or guided or stale" > gpurun_out/r2_t11.log 2>&1
echo "vjp net tests rc=$?"; grep -v "^$" gpurun_out/r2_t11.log | grep "conv_mode\|guided\|passed\|failed\|^E " | head -20
