#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_comb.py -m gpu -q --timeout 60 --timeout-method thread 2>&1 | grep -v "^E  \|^$" | tail -12
