#!/bin/bash
# Round-2 GPU call 3: chained launch as default (+ variants), split foreground kernels, zoom-crop, bench configs, ncu captures.
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1
tail -12 $O/suite.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "default:" "rings33:B200POSE_CHAIN_RINGS=33" "nopipe:B200POSE_FG_PIPELINE=0" "layerwise:B200POSE_CONV_MODE=3" "allold:B200POSE_CONV_MODE=3 B200POSE_FG_PIPELINE=0 B200POSE_LOOKUP_MODE=0 B200POSE_POOL_MODE=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
timeout 200 python tools/conv_counters.py > $O/chain_counters.txt 2>&1; head -16 $O/chain_counters.txt
B200POSE_CHAIN_RINGS=33 timeout 200 python tools/conv_counters.py > $O/chain_counters_rings33.txt 2>&1; head -16 $O/chain_counters_rings33.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 100 --csv --log-file $O/launches_default.csv python tools/profile_step.py --passes 3 > $O/ncu_default.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 5 -c 1 -o $O/chain_ncu -f python tools/profile_step.py --passes 2 > $O/ncu_chain_full.log 2>&1
ncu -i $O/chain_ncu.ncu-rep --page raw --csv > $O/chain_ncu_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fg_weight|fg_target|lm_cluster|corr_lookup_win|geo_to_cl|corr_pool3|flow_head2_partial|context_init" -s 9 -c 9 -o $O/misc_ncu -f python tools/profile_step.py --passes 2 > $O/ncu_misc_full.log 2>&1
ncu -i $O/misc_ncu.ncu-rep --page raw --csv > $O/misc_ncu_raw.csv 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cat $O/bench_cfg1.json; tail -3 $O/bench_cfg1.err
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; cut -c1-600 $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
timeout 600 python bench.py --config cfg1 --global-batch 256 --steps 5 --warmup 3 > $O/bench_cfg1_b256.json 2> $O/bench_cfg1_b256.err; cut -c1-600 $O/bench_cfg1_b256.json; tail -3 $O/bench_cfg1_b256.err
timeout 600 python bench.py --config cfg4 --steps 5 --warmup 3 > $O/bench_cfg4_b8.json 2> $O/bench_cfg4_b8.err; cut -c1-600 $O/bench_cfg4_b8.json; tail -3 $O/bench_cfg4_b8.err
pass=0; fail=0
for i in $(seq 1 5); do
  if timeout 600 python -m pytest tests -m gpu -x -q > $O/suite_loop_$i.txt 2>&1; then pass=$((pass+1)); rm -f $O/suite_loop_$i.txt; else fail=$((fail+1)); fi
done
echo "full-suite loop: $pass passed, $fail failed" | tee $O/suite_loop_summary.txt
ls $O
