#!/bin/bash
# Round-2 GPU call 5: flow head in the chain, chained pre-sums, dynamic unit queue, encoder tests, timings.
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -12 $O/suite.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "default:" "static:B200POSE_CHAIN_DYNAMIC=0" "rings33:B200POSE_CHAIN_RINGS=33" "pipe:B200POSE_FG_PIPELINE=1" "layerwise:B200POSE_CONV_MODE=3"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
timeout 200 python tools/conv_counters.py > $O/chain_counters.txt 2>&1; head -17 $O/chain_counters.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 80 --csv --log-file $O/launches_warm.csv python tools/profile_step.py --passes 3 > $O/ncu_warm.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cat $O/bench_cfg1.json; tail -3 $O/bench_cfg1.err
ls $O
