#!/bin/bash
# Round-2 GPU call 25: per-call set-up halves on two streams (option setup_overlap): tests, A/B step time, bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2y; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_refine.py tests/test_dropin.py -m gpu -q 2>&1 | tail -3 | tee $O/tests.txt
for v in 1 0 1 0; do echo "setup_overlap=$v: $(B200POSE_SETUP_OVERLAP=$v timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt; done
for v in 1 0; do echo "setup_overlap=$v, 8 iterations: $(B200POSE_SETUP_OVERLAP=$v timeout 200 python tools/profile_step.py --passes 2 --time --iters 8 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt; done
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=json.loads([l for l in open('$O/bench_default.json').read().splitlines() if l.startswith('{')][-1]); e=d['e2e']; print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', e['value'], e['context'], 'oracle', d['oracle_check'])"
