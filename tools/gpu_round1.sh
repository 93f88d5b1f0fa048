#!/bin/bash
# One GPU session: second-generation conv kernel validation, timing per mode, full GPU suite, launch list, ncu capture.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
echo "== conv second generation tests" > gpurun_out/r.log
timeout 600 python -m pytest tests/test_gpu_conv_layers.py -q -k second_generation --tb=line 2>&1 | tail -40 >> gpurun_out/r.log
echo "== timing per mode" >> gpurun_out/r.log
for m in 0 1 2 3; do
  echo "-- mode $m" >> gpurun_out/r.log
  B200POSE_CONV_MODE=$m timeout 300 python tools/profile_step.py --time --passes 2 2>&1 | tail -3 >> gpurun_out/r.log
done
echo "== full gpu suite" >> gpurun_out/r.log
timeout 700 python -m pytest tests -q -m gpu -x -k "not second_generation" --durations=8 2>&1 | tail -25 >> gpurun_out/r.log
echo "== launch list mode 3" >> gpurun_out/r.log
B200POSE_CONV_MODE=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 110 --csv --log-file gpurun_out/launches_mode3.csv python tools/profile_step.py --passes 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log >> gpurun_out/r.log
echo "== ncu full conv_umma2 mode 3" >> gpurun_out/r.log
B200POSE_CONV_MODE=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 5 -c 11 -o gpurun_out/conv_umma2_mode3 -f python tools/profile_step.py --passes 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log >> gpurun_out/r.log
echo "== bench mode 0 / mode 3" >> gpurun_out/r.log
B200POSE_CONV_MODE=0 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_mode0.json 2> gpurun_out/bench_mode0.err
B200POSE_CONV_MODE=3 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_mode3.json 2> gpurun_out/bench_mode3.err
cat gpurun_out/bench_mode0.json gpurun_out/bench_mode3.json >> gpurun_out/r.log
cat gpurun_out/r.log
