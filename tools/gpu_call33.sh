#!/bin/bash
# Round-2 GPU call 33: configs[3] and the B=256 strong-scaling point with the final defaults.
cd "$(dirname "$0")/.."
O=gpurun_out/r3h; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --config cfg1 --global-batch 256 --steps 5 --warmup 3 > $O/bench_cfg1_b256.json 2> $O/bench_cfg1_b256.err
python - <<'PY'
import json
for f in ('bench_cfg3', 'bench_cfg1_b256'):
    d = json.loads([l for l in open(f'gpurun_out/r3h/{f}.json').read().splitlines() if l.startswith('{')][-1]); e = d['e2e']
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(e['value']), 'frac', round(d['roofline']['frac'], 4))
PY
