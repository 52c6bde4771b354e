#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_vjp.py -m gpu -x -q -s --timeout 900 --timeout-method thread 2>&1 | grep -i "rel\|passed\|failed\|error\|gradient" | tail -20
timeout 300 python tools/bench_vjp.py 2>/dev/null | tail -1
AID_VJP_TC1X1=0 timeout 300 python tools/bench_vjp.py 2>/dev/null | tail -1
timeout 300 python tools/bench_vjp.py --net paper_44k --len 184184 2>/dev/null | tail -1
