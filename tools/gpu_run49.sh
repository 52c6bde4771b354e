#!/bin/bash
# fused init block: block-level parity + timing, bench-shape invariants, compute-sanitizer on the new kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_init_block.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -40
timeout 600 python -m pytest tests/test_gpu_bench_shape.py -m gpu -x -q -s 2>&1 | grep "whole batch\|row \|passed\|failed\|Error\|assert" | head -30
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_init_block.py -m gpu -q -k "case0 or case2 or case5" 2>&1 | grep "ERROR SUMMARY\|passed\|failed\|Invalid" | head
