#!/bin/bash
# Round-2 GPU call 11: host-side gather of the context texels (e2e), immediate-offset upsample kernel; suite + bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2k; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
lscpu > $O/lscpu.txt 2>&1; grep -i "model name\|^CPU(s)\|NUMA node" $O/lscpu.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -6 $O/suite.txt
timeout 300 python tools/e2e_sweep.py --threads -1,1,2,4,6,8,12,16 > $O/e2e_sweep.txt 2>&1; cat $O/e2e_sweep.txt
timeout 200 python tools/profile_step.py --passes 2 --time > $O/time_default.txt 2>&1; grep 'ms per pass' $O/time_default.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_step.csv python tools/profile_step.py --passes 1 > $O/ncu_step.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2k/launches_step.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; kn = H.index('Kernel Name'); mv = H.index('Metric Value')
acc = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) <= mv: continue
    n = r[kn].split('(')[0][:60]; v = float(r[mv].replace(',', '')) / 1e3
    a = acc.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
for n, (c, t) in acc.items(): print(f"{n:62s} x{c:3d} {t / c:8.1f} us each")
PY
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cut -c1-300 $O/bench_cfg1.json; tail -2 $O/bench_cfg1.err
python -c "
import json; d=json.load(open('$O/bench_cfg1.json')); e=d['e2e']; print('e2e', e['value'], e['context']); [print('  ', r) for r in e['variants']]; print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])"
ls $O
