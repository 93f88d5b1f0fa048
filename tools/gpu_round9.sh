#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r9.log
: > $L
for m in 0 3; do B200POSE_CONV_MODE=$m timeout 200 python tools/conv_counters.py 2>&1 | tail -34 >> $L; done
cat $L
