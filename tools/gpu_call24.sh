#!/bin/bash
# Round-2 GPU call 24: HEAD after the upsample restructuring: full suite, step time, kernel time, bench default.
cd "$(dirname "$0")/.."
O=gpurun_out/r2x; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -4 $O/suite.txt
timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass' | tee $O/time_default.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample_weight -c 8 --csv --log-file $O/upw.csv python tools/profile_step.py --passes 2 > $O/upw.log 2>&1
echo "upsample default: $(grep upsample_weight $O/upw.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=json.loads([l for l in open('$O/bench_default.json').read().splitlines() if l.startswith('{')][-1]); e=d['e2e']; print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', e['value'], e['context'], 'frac', d['roofline']['frac'], 'enc', d['encoder']['ms_per_batch'], d['clocks'])"
