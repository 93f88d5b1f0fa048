#!/bin/bash
# Round-2 GPU call 22: upsample kernel after restoring the plane-pointer load form; windowed-copy test; step time.
cd "$(dirname "$0")/.."
O=gpurun_out/r2v; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample_weight -c 8 --csv --log-file $O/upw.csv python tools/profile_step.py --passes 2 > $O/upw.log 2>&1
echo "upsample default: $(grep upsample_weight $O/upw.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass' | tee $O/time_default.txt
timeout 600 python -m pytest tests/test_gpu_refine.py tests/test_gpu_ops.py tests/test_gpu_encoder.py -m gpu -q 2>&1 | tail -3 | tee $O/tests.txt
