#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/r12.log
echo "== tests" > $L
timeout 400 python -m pytest tests/test_gpu_refine.py tests/test_gpu_ops.py -q -x 2>&1 | tail -3 >> $L
run() { echo "-- $*" >> $L; env "$@" timeout 200 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; }
echo "== step timing" >> $L
run B200POSE_FG_LIST=0
run B200POSE_FG_UPSAMPLE=0
run B200POSE_FG_UPSAMPLE=1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 70 -c 120 --csv --log-file gpurun_out/launches_fg.csv python tools/profile_step.py --passes 3 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_fg.csv 2>/dev/null | sed -n 14,20p >> $L
cat $L
