#!/bin/bash
# Round-2 GPU call 14: suite on the new defaults; encoder after the stem / statistics changes; step time.
cd "$(dirname "$0")/.."
O=gpurun_out/r2n; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -5 $O/suite.txt
timeout 300 python tools/profile_encoder.py --time > $O/encoder_time.txt 2>&1; cat $O/encoder_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/encoder_launches.csv python tools/profile_encoder.py --passes 1 > $O/encoder_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2n/encoder_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; kn = H.index('Kernel Name'); mv = H.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv or 'at::' in r[kn] or 'pack' in r[kn]: continue
    a = agg.setdefault(r[kn].split('(')[0][-40:], [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(',', '')) / 1e3
for k, v in agg.items(): print(f"{k:42s} x{v[0]:3d} {v[1]:8.1f} us")
PY
timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass' | tee $O/time_default.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample_weight -c 8 --csv --log-file $O/upw.csv python tools/profile_step.py --passes 2 > $O/upw.log 2>&1
echo "upsample default: $(grep upsample_weight $O/upw.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
ls $O
