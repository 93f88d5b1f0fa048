#!/bin/bash
# Evidence run: bench line, ncu launch list of the same workload, ncu --set full of the convolution kernels.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 120 --csv --log-file gpurun_out/launches_ncu.csv python tools/profile_step.py --passes 3 > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 5 -c 11 -o gpurun_out/conv_ncu -f python tools/profile_step.py --passes 1 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/conv_ncu.ncu-rep --page raw --csv > gpurun_out/conv_ncu_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"upsample_weight|corr_lookup|flow_head2|lm_multi|context_init|tiled_convert" -s 2 -c 8 -o gpurun_out/misc_ncu -f python tools/profile_step.py --passes 1 > gpurun_out/ncu_misc.log 2>&1
ncu -i gpurun_out/misc_ncu.ncu-rep --page raw --csv > gpurun_out/misc_ncu_raw.csv 2>/dev/null
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
