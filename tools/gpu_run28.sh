#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_comb.py -m gpu -q --timeout 60 --timeout-method thread 2>&1 | grep -v "^E  \|^$" | tail -8
AID_COMB_DEBUG=1 timeout 200 python tools/time_comb.py 2>&1 | grep -v "comb transform\|comb mma" | tail -16
