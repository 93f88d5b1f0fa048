#!/bin/bash
# Round-2 GPU call 7: x-major horizontal-tap reuse for the 1x5 layers of the chained launch.
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 600 python -m pytest tests/test_gpu_refine.py tests/test_gpu_ops.py -q ) > $O/tests.txt 2>&1; tail -8 $O/tests.txt
T="timeout 200 python tools/profile_step.py --passes 2 --time"
for cfg in "xmajor:" "ymajor:B200POSE_CHAIN_XMAJOR=0" "ymajor_r33:B200POSE_CHAIN_XMAJOR=0 B200POSE_CHAIN_RINGS=33"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs $T > $O/time_$name.txt 2>&1; echo "$name: $(grep 'ms per pass' $O/time_$name.txt)"
done
timeout 200 python tools/conv_counters.py > $O/chain_counters_xmajor.txt 2>&1; sed -n 1,19p $O/chain_counters_xmajor.txt
ls $O
