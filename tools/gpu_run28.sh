#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_comb.py -m gpu -x -q --timeout 60 --timeout-method thread 2>&1 | tail -4
AID_COMB_DEBUG=1 timeout 200 python tools/time_comb.py 2>&1 | tail -24
