#!/bin/bash
# Round-2 GPU call 28: HEAD: full GPU suite, smoke, default bench (the driver's sequence).
cd "$(dirname "$0")/.."
O=gpurun_out/r3g; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/suite.txt 2>&1; tail -4 $O/suite.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=json.loads([l for l in open('$O/bench_default.json').read().splitlines() if l.startswith('{')][-1]); e=d['e2e']; print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', e['value'], e['context'], 'frac', d['roofline']['frac'], 'enc', d['encoder']['ms_per_batch'], 'launches', d['gpu_launches'])"
