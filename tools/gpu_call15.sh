#!/bin/bash
# Round-2 GPU call 15: the driver's sequence on the committed defaults: smoke, bench (both arms), cfg3.
cd "$(dirname "$0")/.."
O=gpurun_out/r2o; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-200 $O/bench_reference.json
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json')); e=d['e2e']; print('value', d['value'], 'ms', d['ms_per_step'], 'launches', d['gpu_launches']); print('e2e', e['value'], e['context']); [print('  ', r) for r in e['variants']]; print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch']); print('encoder', d['encoder']); print('oracle', d['oracle_check'], d['clocks'])"
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python -c "
import json; d=json.load(open('$O/bench_cfg3.json')); e=d['e2e']; print('cfg3 value', d['value'], 'e2e', e['value'], e['context'])"
ls $O
