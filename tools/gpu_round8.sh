#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r8.log
echo "== conv tests transposed" > $L
timeout 400 python -m pytest tests/test_gpu_conv_layers.py -q --tb=line -k "second_generation and transposed" 2>&1 | tail -12 >> $L
echo "== refine tests" >> $L
timeout 400 python -m pytest tests/test_gpu_refine.py -q -x -s -k "transposed" 2>&1 | grep "parity\|passed\|failed\|rror" | tail >> $L
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_CONV_MODE=3
run B200POSE_CONV_MODE=8
run B200POSE_CONV_MODE=11
B200POSE_CONV_MODE=11 timeout 200 python tools/conv_counters.py >> $L 2>&1
cat $L
