#!/bin/bash
# final validation of the round: GPU suite, smoke, bench line, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/r2_bench_v4.err | tail -1 | tee gpurun_out/r2_bench_tc2_v4.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_tc2_b8_v3.csv python tools/profile_forward.py --batch 8 > gpurun_out/r2_ncu_ll.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_tc2_b8_v3.csv "one forward, B=8 x 262144, conv_mode 2 (end of round 2: fused init / out blocks, in-conversion upsampling)" > gpurun_out/r2_launches_tc2_b8_v3.summary.txt; cat gpurun_out/r2_launches_tc2_b8_v3.summary.txt
