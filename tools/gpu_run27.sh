#!/bin/bash
timeout 300 python tools/time_gn.py 2>&1 | tail -7
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_bench_shape.py -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -3
