#!/bin/bash
# Round-2 GPU call 4: encoder (f2) first run, zoom-crop, pipeline A/B at 8 iterations, warm launch list, bench with kernel-only roofline.
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 600 python -m pytest tests/test_gpu_encoder.py -q -x ) > $O/encoder_tests.txt 2>&1; tail -25 $O/encoder_tests.txt
( time timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_encoder.py ) > $O/suite.txt 2>&1; tail -8 $O/suite.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "pipe_i4:" "nopipe_i4:B200POSE_FG_PIPELINE=0" ; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
for cfg in "pipe_i8:" "nopipe_i8:B200POSE_FG_PIPELINE=0" ; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T --iters 8 > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
# warm-cache launch lists (no cache flush between kernels): closer to the real step than the default cold-cache list
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 40 -c 100 --csv --log-file $O/launches_warm_pipe.csv python tools/profile_step.py --passes 3 > $O/ncu_warm_pipe.log 2>&1
B200POSE_FG_PIPELINE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 40 -c 100 --csv --log-file $O/launches_warm_nopipe.csv python tools/profile_step.py --passes 3 > $O/ncu_warm_nopipe.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cat $O/bench_cfg1.json; tail -3 $O/bench_cfg1.err
ls $O
