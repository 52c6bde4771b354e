#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vjp.py -m gpu -q -s -k "denoiser_vjp or guided or stale" > gpurun_out/r2_t13.log 2>&1
echo "vjp net tests rc=$?"; grep -v "^$" gpurun_out/r2_t13.log | grep "conv_mode\|guided\|passed\|failed\|^E " | head -20
python tools/bench_vjp.py --len 262144 > gpurun_out/r2_vjp_22k.json 2> gpurun_out/r2_vjp.err
python tools/bench_vjp.py --net paper_44k --len 184184 > gpurun_out/r2_vjp_44k.json 2>> gpurun_out/r2_vjp.err
cat gpurun_out/r2_vjp_22k.json gpurun_out/r2_vjp_44k.json; tail -3 gpurun_out/r2_vjp.err
