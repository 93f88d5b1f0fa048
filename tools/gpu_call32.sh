#!/bin/bash
# Round-2 GPU call 32: which loop kernels should keep the PDL attribute?  (option pdl_off bit mask; lookup_mode 2 = plain launch)
cd "$(dirname "$0")/.."
O=gpurun_out/r3f; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
for m in 0 1 4 16 32 48 8 53; do
  echo "pdl_off=$m: $(B200POSE_PDL_OFF=$m timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt
done
for m in 0 48; do
  echo "B=1 pdl_off=$m: $(B200POSE_PDL_OFF=$m timeout 200 python tools/profile_step.py --passes 2 --time --batch 1 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt
done
echo "B=1 lookup_mode=1: $(B200POSE_LOOKUP_MODE=1 timeout 200 python tools/profile_step.py --passes 2 --time --batch 1 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt
