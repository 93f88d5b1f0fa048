#!/bin/bash
# Experiments: where does the second-generation conv kernel lose time in the pipelined step?
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r2.log
echo "== update_block / refine tests with the new flow head" > $L
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_refine.py -q -x -k "update or golden or batched" 2>&1 | tail -4 >> $L
echo "== per-layer bench" >> $L
timeout 300 python tools/conv_layer_bench.py >> $L 2>&1
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_CONV_MODE=0
run B200POSE_CONV_MODE=4
run B200POSE_CONV_MODE=3
run B200POSE_CONV_MODE=3 B200POSE_V2_NOPDL=1
run B200POSE_CONV_MODE=3 B200POSE_V2_BUDGET_KB=190
run B200POSE_CONV_MODE=3 B200POSE_CONV_LAYERS=2      # C2 only
run B200POSE_CONV_MODE=3 B200POSE_CONV_LAYERS=512    # HEADS only
run B200POSE_CONV_MODE=3 B200POSE_CONV_LAYERS=386    # C2, ZR2, Q2
run B200POSE_CONV_MODE=2 B200POSE_CONV_LAYERS=386
cat $L
