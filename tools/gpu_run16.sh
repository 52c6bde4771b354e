#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_t16_all.log 2>&1
echo "all tests rc=$?"; tail -16 gpurun_out/r2_t16_all.log
grep -h "rel-L2 per block at\|vs reference golden\|row .* of the batch" gpurun_out/r2_t16_all.log | head
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
python -c "
import json; d = json.load(open('gpurun_out/r2_bench5.json')); print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['kernel_share_of_step'], d['clocks'], d.get('fp32_grade', {}).get('value'), d['cpu_baseline'], d['gpu_launches'])"
