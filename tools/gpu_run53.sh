#!/bin/bash
# launch list of one forward at B=8 after the init / out block fusions; ncu --set full of the two new kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_tc2_b8_v2.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_ll.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_tc2_b8_v2.csv "one forward, B=8 x 262144, conv_mode 2 (round 2, with fused init and out blocks)" > gpurun_out/r2_launches_tc2_b8_v2.summary.txt; cat gpurun_out/r2_launches_tc2_b8_v2.summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"init_block_kernel|out_block_kernel" -c 6 -f -o gpurun_out/r2_ncu_init_out_full python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_f3.log 2>&1
echo "ncu full rc=$?"
