#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_comb_kernel -c 2 -f -o gpurun_out/r2_ncu_conv_comb_full python tools/time_comb_one.py 1 > gpurun_out/r2_ncu_comb.log 2>&1
echo "ncu comb rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -c 2 -f -o gpurun_out/r2_ncu_conv_tc2_l0_full python tools/time_comb_one.py 0 > gpurun_out/r2_ncu_tc2l0.log 2>&1
echo "ncu tc2 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
