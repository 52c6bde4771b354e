#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_tc.py -m gpu -q -s -k "rows_equal_solo or per_clip_sigma_and_full_size" 2>&1 | grep "row \|run-to-run\|passed\|failed"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_vjp_b1.csv python tools/profile_vjp.py --batch 1 > gpurun_out/r2_ncu_vjp.log 2>&1
echo "vjp launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_vjp_b1.csv "backward (VJP) of one clip x 262144, conv_mode 2" > gpurun_out/r2_launches_vjp_b1.summary.txt; cat gpurun_out/r2_launches_vjp_b1.summary.txt
