#!/bin/bash
# Round-2 GPU call 23: A/B of the dense upsample kernel on ONE box: the tree at commit c2b3c29 (build_ab/old) vs HEAD.
cd "$(dirname "$0")/.."
O=gpurun_out/r2w; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
for side in old new old new; do
  if [ $side = old ]; then D=build_ab/old; else D=.; fi
  ( cd $D && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample_weight -c 8 --csv --log-file /tmp/upw_$side.csv python tools/profile_step.py --passes 2 > /tmp/upw_$side.log 2>&1 )
  echo "$side: $(grep upsample_weight /tmp/upw_$side.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')" | tee -a $O/ab.txt
done
