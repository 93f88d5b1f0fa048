#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r10.log
echo "== gpu suite" > $L
timeout 700 python -m pytest tests -q -m gpu -x -s 2>&1 | grep "parity\|passed\|failed\|rror" | tail -30 >> $L
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_CONV_MODE=0
run B200POSE_CONV_MODE=3
B200POSE_CONV_MODE=3 timeout 200 python tools/conv_counters.py 2>&1 | head -16 >> $L
cat $L
