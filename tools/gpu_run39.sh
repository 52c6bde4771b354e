#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread 2>&1 | tail -3
timeout 300 python tools/bench_vjp.py 2>/dev/null | tail -1
timeout 300 python tools/time_forward.py 1 2 4 8 2>/dev/null | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['fused_layers']['frac'], d['clocks'])"
