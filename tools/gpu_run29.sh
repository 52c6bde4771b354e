#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_comb.py tests/test_gpu_bench_shape.py tests/test_gpu_unet.py tests/test_gpu_tc.py -m gpu -x -q --timeout 600 --timeout-method thread 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_comb.json 2> gpurun_out/r2_bench_comb.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_comb.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], {k:d['roofline'][k] for k in ('achieved','frac','kernel_ms_per_step','kernel_share_of_step')}, d['clocks'])
PY
