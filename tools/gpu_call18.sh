#!/bin/bash
# Round-2 GPU call 18: windowed copy of the second descriptor map in the host entry: tests, e2e sweep over margins, bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2r; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests/test_gpu_refine.py tests/test_gpu_ops.py -m gpu -q -x ) > $O/tests.txt 2>&1; tail -5 $O/tests.txt
for cfg in "sparse_off:B200POSE_SPARSE_G2=0" "margin24:B200POSE_G2_MARGIN=24" "margin16:B200POSE_G2_MARGIN=16" "margin40:B200POSE_G2_MARGIN=40"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/e2e_sweep.py --threads=-1,8 > $O/e2e_$name.txt 2>&1; echo "== $name"; grep "host entry" $O/e2e_$name.txt
done
timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass' | tee $O/time_default.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json')); e=d['e2e']; print('value', d['value'], 'e2e', e['value'], e['context']); [print('  ', r) for r in e['variants']]"
