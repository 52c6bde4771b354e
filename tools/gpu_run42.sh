#!/bin/bash
timeout 200 python tools/time_comb.py 2>&1 | grep -v "comb transform\|comb mma\|comb96 mma" | grep "fused 1" | tail -8
timeout 200 python tools/time_conv.py 3 5x3 2>&1 | tail -6
