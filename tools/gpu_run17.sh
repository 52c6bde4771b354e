#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_tc2_b8.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_ll.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_tc2_b8.csv "one forward, B=8 x 262144, conv_mode 2 (round 2 final)" > gpurun_out/r2_launches_tc2_b8.summary.txt; cat gpurun_out/r2_launches_tc2_b8.summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2_kernel -s 40 -c 3 -f -o gpurun_out/r2_ncu_conv_tc2_full python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_f1.log 2>&1
echo "ncu full conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc_kernel -c 2 -f -o gpurun_out/r2_ncu_attention_tc_full python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_f2.log 2>&1
echo "ncu full attention rc=$?"
