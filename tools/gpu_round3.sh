#!/bin/bash
# What bounds the convolution kernels: drop one of TMA / MMA / epilogue (results garbage, timing only).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r3.log
: > $L
for dbg in 0 1 2 4 3 6 8; do
  echo "== B200POSE_V2_DEBUG=$dbg (1 no TMA, 2 no MMA, 4 no epilogue, 8 extra commit [gen1 only])" >> $L
  B200POSE_V2_DEBUG=$dbg timeout 200 python tools/conv_layer_bench.py --modes 0,4,3 --layers 0,1,5,6,7,9 >> $L 2>&1
done
cat $L
