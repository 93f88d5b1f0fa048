#!/bin/bash
# Round-2 GPU call 9: interleaved segments, cluster LM as default, LM backward; timings + counters; bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -8 $O/suite.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "default:" "nomerge:B200POSE_CHAIN_MERGE=0" "dynamic:B200POSE_CHAIN_DYNAMIC=1" "lm_spin:B200POSE_LM_CLUSTER=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
timeout 200 python tools/conv_counters.py > $O/chain_counters.txt 2>&1; sed -n 1,19p $O/chain_counters.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cut -c1-300 $O/bench_cfg1.json; tail -2 $O/bench_cfg1.err
python -c "
import json; d=json.load(open('$O/bench_cfg1.json')); print('encoder leg', d.get('encoder')); print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])"
ls $O
