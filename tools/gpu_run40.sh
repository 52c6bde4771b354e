#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_comb.py tests/test_gpu_bench_shape.py -m gpu -x -q -s --timeout 900 --timeout-method thread 2>&1 | grep "row \|fused layers vs\|passed\|failed\|Error" | tail -14
AID_COMB_DEBUG=1 timeout 200 python tools/time_comb.py 2>&1 | grep -v "comb transform\|comb mma" | tail -12
