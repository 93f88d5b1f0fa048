#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r5.log
echo "== conv tests second generation" > $L
timeout 400 python -m pytest tests/test_gpu_conv_layers.py -q --tb=line -k second_generation 2>&1 | tail -5 >> $L
echo "== per-layer bench (N256 pairs)" >> $L
timeout 300 python tools/conv_layer_bench.py --modes 0,1,3 >> $L 2>&1
echo "== per-layer bench (N128 pairs)" >> $L
B200POSE_PAIR_N256=0 timeout 300 python tools/conv_layer_bench.py --modes 1,3 --layers 0,5,7,9 >> $L 2>&1
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_CONV_MODE=0
run B200POSE_CONV_MODE=1
run B200POSE_CONV_MODE=3
cat $L
