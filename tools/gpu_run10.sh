#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vjp.py -m gpu -q -k "not denoiser_vjp and not guided" > gpurun_out/r2_t10.log 2>&1
echo "vjp op tests rc=$?"; tail -8 gpurun_out/r2_t10.log
python tools/dbg_vjp.py 2>&1 | grep -v "Attention layer\|^$" | tail -20
