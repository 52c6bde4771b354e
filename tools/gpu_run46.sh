#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vjp.py -m gpu -x -q --timeout 900 --timeout-method thread 2>&1 | tail -4
timeout 300 python tools/bench_vjp.py 2>/dev/null | tail -1
timeout 300 python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 35 --xi 0.25 2>/dev/null | tail -1
