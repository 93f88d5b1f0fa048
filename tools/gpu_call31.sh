#!/bin/bash
# Round-2 GPU call 31: why is the faster lookup kernel slower in the live step?  Isolated live timing + block-size / PDL builds.
cd "$(dirname "$0")/.."
O=gpurun_out/r3e; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 200 python tools/lookup_live.py 2>&1 | tee $O/lookup_live.txt
for m in 1 2 4 5 6; do
  echo "lookup_mode=$m: $(B200POSE_LOOKUP_MODE=$m timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt
done
