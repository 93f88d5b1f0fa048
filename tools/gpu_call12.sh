#!/bin/bash
# Round-2 GPU call 12: e2e sweep over host-gather threads / modes; dense upsample kernel variants (ncu durations + step time).
cd "$(dirname "$0")/.."
O=gpurun_out/r2l; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
for m in 0 1 2 3; do
  B200POSE_HOST_GATHER=$m timeout 300 python tools/e2e_sweep.py --threads=-1,2,4,8,12,16 > $O/e2e_sweep_mode$m.txt 2>&1; echo "== host_gather=$m"; cat $O/e2e_sweep_mode$m.txt
done
for v in 0 1 2 3 4; do
  B200POSE_UPSAMPLE_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample_weight -c 8 --csv --log-file $O/upw_v$v.csv python tools/profile_step.py --passes 2 > $O/upw_v$v.log 2>&1
  echo "variant $v: $(grep upsample_weight $O/upw_v$v.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
  B200POSE_UPSAMPLE_VARIANT=$v timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass'
done
for v in 0 3; do
  B200POSE_UPSAMPLE_VARIANT=$v timeout 300 ncu --set full --clock-control none -k regex:upsample_weight -s 1 -c 1 --csv --page raw --log-file $O/upw_full_v$v.csv python tools/profile_step.py --passes 1 > $O/upw_full_v$v.log 2>&1
done
ls $O
