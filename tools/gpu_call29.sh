#!/bin/bash
# Round-2 GPU call 29: row-per-thread lookup kernel (lookup_mode 2): tests, kernel time, step time.
cd "$(dirname "$0")/.."
O=gpurun_out/r3c; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "lookup or corr" 2>&1 | tail -3 | tee $O/tests.txt
for m in 1 2; do
  B200POSE_LOOKUP_MODE=$m timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:corr_lookup -c 4 --csv --log-file $O/lk_$m.csv python tools/profile_step.py --passes 1 > $O/lk_$m.log 2>&1
  echo "lookup_mode=$m: $(grep corr_lookup $O/lk_$m.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tr '\n' ' ')" | tee -a $O/ab.txt
  echo "lookup_mode=$m: $(B200POSE_LOOKUP_MODE=$m timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt
done
