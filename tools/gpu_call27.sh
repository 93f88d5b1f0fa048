#!/bin/bash
# Round-2 GPU call 27: partial host gather (option host_gather_planes): tests and the bench's three e2e variants at N=1.
cd "$(dirname "$0")/.."
O=gpurun_out/r3a; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_refine.py tests/test_gpu_ops.py -m gpu -q -k "host or context or texel" 2>&1 | tail -3 | tee $O/tests.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python -c "
import json; d=json.loads([l for l in open('$O/bench_default.json').read().splitlines() if l.startswith('{')][-1]); e=d['e2e']; print('value', d['value'], 'e2e', e['value'], e['context']); [print('  ', r['context'], round(r['value']), round(r['ms_per_step'],2), round(r['h2d_gbs_this_rank'],1)) for r in e['variants']]"
