#!/bin/bash
# memcheck of the fused kernels on two small shapes (out-of-bounds global / shared accesses)
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_comb.py -m gpu -x -q --timeout 500 --timeout-method thread -k "case1 or case8" 2>&1 | tail -12
