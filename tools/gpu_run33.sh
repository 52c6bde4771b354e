#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_unet.py tests/test_gpu_bench_shape.py tests/test_gpu_sampler.py -m gpu -x -q --timeout 900 --timeout-method thread 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench:', d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
AID_ATT_FOLD=0 timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no fold:', d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
