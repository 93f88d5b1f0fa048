#!/bin/bash
# Round-2 GPU call 6: dynamic queue with prefetch vs static, encoder + drop-in tests, per-cluster balance, bench cfg3/cfg4 sweep.
cd "$(dirname "$0")/.."
O=gpurun_out/r2f; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -8 $O/suite.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "static:" "dynamic:B200POSE_CHAIN_DYNAMIC=1" "static_r33:B200POSE_CHAIN_RINGS=33" "dynamic_r33:B200POSE_CHAIN_DYNAMIC=1 B200POSE_CHAIN_RINGS=33"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
timeout 200 python tools/conv_counters.py > $O/chain_counters_static.txt 2>&1; sed -n 1,18p $O/chain_counters_static.txt
B200POSE_CHAIN_DYNAMIC=1 timeout 200 python tools/conv_counters.py > $O/chain_counters_dynamic.txt 2>&1; sed -n 14,18p $O/chain_counters_dynamic.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cut -c1-300 $O/bench_cfg1.json; tail -2 $O/bench_cfg1.err
timeout 900 python bench.py --config cfg4 --sweep --steps 3 --warmup 3 > $O/bench_cfg4_sweep.json 2> $O/bench_cfg4_sweep.err; python -c "
import json; d=json.load(open('$O/bench_cfg4_sweep.json')); [print(r) for r in d['sweep']]"; tail -2 $O/bench_cfg4_sweep.err
ls $O
