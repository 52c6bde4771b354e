#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --timeout 90 --timeout-method thread -k "single_fp16 or cta_pair or single_cta or 1x1" 2>&1 | tail -3
timeout 120 python tools/time_conv.py 3 5x3 2>&1 | tail -6
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches_tc2_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2b_ncu_ll.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2b_launches_tc2_b8.csv "one forward, B=8 x 262144, conv_mode 2 (static epilogue, cta_group::2 on the 256 / 96-channel layers)" > gpurun_out/r2b_launches_tc2_b8.summary.txt; cat gpurun_out/r2b_launches_tc2_b8.summary.txt
