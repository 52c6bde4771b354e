#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_comb.py tests/test_gpu_bench_shape.py tests/test_gpu_unet.py tests/test_gpu_tc.py -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -4
for v in 1 0; do
AID_COMB96=$v timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AID_COMB96=$v:', d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['clocks'])"
done
AID_COMB96=1 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AID_COMB96=1 again:', d['value'], d['roofline']['frac'], d['clocks'])"
