#!/bin/bash
# Round-2 GPU call 2: new kernels (window lookup, pool3, foreground pipeline + cluster LM), chain counters, A/B timings, launch lists.
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time B200POSE_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1
tail -15 $O/suite.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "default:" "nopipe:B200POSE_FG_PIPELINE=0" "oldlookup:B200POSE_LOOKUP_MODE=0" "oldpool:B200POSE_POOL_MODE=0" "chain:B200POSE_CONV_MODE=19" "lm_noacc:B200POSE_LM_DEBUG=1" "allold:B200POSE_FG_PIPELINE=0 B200POSE_LOOKUP_MODE=0 B200POSE_POOL_MODE=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(tail -2 $O/time_$name.txt | head -1)"
done
B200POSE_CONV_MODE=19 timeout 200 python tools/conv_counters.py > $O/chain_counters.txt 2>&1; head -20 $O/chain_counters.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 140 --csv --log-file $O/launches_default.csv python tools/profile_step.py --passes 3 > $O/ncu_default.log 2>&1
B200POSE_CONV_MODE=19 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 100 --csv --log-file $O/launches_chain.csv python tools/profile_step.py --passes 3 > $O/ncu_chain.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json
pass=0; fail=0
for i in $(seq 1 5); do
  if timeout 600 python -m pytest tests -m gpu -x -q > $O/suite_loop_$i.txt 2>&1; then pass=$((pass+1)); rm -f $O/suite_loop_$i.txt; else fail=$((fail+1)); fi
done
echo "full-suite loop: $pass passed, $fail failed" | tee $O/suite_loop_summary.txt
