#!/bin/bash
# round-2 GPU session 1: new parity tests, bench line, ncu evidence for the narrow-N conv layers and the operand pass
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_device_sampler.py "tests/test_gpu_unet.py" -m gpu -q -x -s --durations=10 > gpurun_out/r2_t1_new.log 2>&1
echo "new tests rc=$?" 
tail -5 gpurun_out/r2_t1_new.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench1.json
python tools/time_conv.py 3 5x3 > gpurun_out/r2_time_conv.log 2>&1
python tools/time_gn.py >> gpurun_out/r2_time_conv.log 2>&1
cat gpurun_out/r2_time_conv.log
TC_SHAPES="8,64,64,4096,2;8,96,192,1024,4;8,128,256,512,16" timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -c 6 -f -o gpurun_out/r2_ncu_conv_tc2_narrow python tools/time_conv.py 3 5x3 > gpurun_out/r2_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_act_tc2_kernel -s 2 -c 4 -f -o gpurun_out/r2_ncu_gn_act_tc2 python tools/time_gn.py > gpurun_out/r2_ncu2.log 2>&1
echo "ncu2 rc=$?"
python tools/bench_sampler.py --config inpaint --batch 8 --gap-ms 300 --steps 8 > gpurun_out/r2_samp_dev.json 2> gpurun_out/r2_samp_dev.err
python tools/bench_sampler.py --config inpaint --batch 8 --gap-ms 300 --steps 8 --noise host > gpurun_out/r2_samp_host.json 2>> gpurun_out/r2_samp_dev.err
python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 8 > gpurun_out/r2_samp_dev_b1.json 2>> gpurun_out/r2_samp_dev.err
python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 8 --no-graph > gpurun_out/r2_samp_dev_b1_nograph.json 2>> gpurun_out/r2_samp_dev.err
python tools/bench_sampler.py --config inpaint --batch 1 --gap-ms 300 --steps 8 --noise host > gpurun_out/r2_samp_host_b1.json 2>> gpurun_out/r2_samp_dev.err
cat gpurun_out/r2_samp_*.json; tail -5 gpurun_out/r2_samp_dev.err
