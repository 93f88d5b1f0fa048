#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r4.log
echo "== conv tests (all generations)" > $L
timeout 400 python -m pytest tests/test_gpu_conv_layers.py -q --tb=line 2>&1 | tail -8 >> $L
echo "== per-layer bench" >> $L
timeout 300 python tools/conv_layer_bench.py --modes 0,4,1,2,3 >> $L 2>&1
echo "== MMA only (debug 1)" >> $L
B200POSE_V2_DEBUG=1 timeout 300 python tools/conv_layer_bench.py --modes 0,4,3 --layers 1,5,6,9 >> $L 2>&1
echo "== TMA only (debug 6)" >> $L
B200POSE_V2_DEBUG=6 timeout 300 python tools/conv_layer_bench.py --modes 0,4,3 --layers 1,5,6,9 >> $L 2>&1
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_CONV_MODE=0
run B200POSE_CONV_MODE=1
run B200POSE_CONV_MODE=2
run B200POSE_CONV_MODE=3
run B200POSE_CONV_MODE=4
cat $L
