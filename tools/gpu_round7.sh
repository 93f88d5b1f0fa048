#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r7.log
echo "== gpu suite" > $L
timeout 700 python -m pytest tests -q -m gpu -x -s 2>&1 | grep "parity\|passed\|failed\|Error\|error" | tail -30 >> $L
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_CONV_MODE=0
run B200POSE_CONV_MODE=1
run B200POSE_CONV_MODE=3
run B200POSE_CONV_MODE=3 B200POSE_PAIR_N256=0
for m in 0 3; do B200POSE_CONV_MODE=$m timeout 200 python tools/conv_counters.py >> $L 2>&1; done
cat $L
