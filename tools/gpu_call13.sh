#!/bin/bash
# Round-2 GPU call 13: encoder launch list; upsample variants 5/6; chain DRAM traffic with warm caches; suite on new defaults.
cd "$(dirname "$0")/.."
O=gpurun_out/r2m; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python tools/profile_encoder.py --time > $O/encoder_time.txt 2>&1; cat $O/encoder_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $O/encoder_launches.csv python tools/profile_encoder.py --passes 1 > $O/encoder_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2m/encoder_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; kn = H.index('Kernel Name'); mn = H.index('Metric Name'); mv = H.index('Metric Value'); idc = H.index('ID')
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    d = per.setdefault(r[idc], {'name': r[kn].split('(')[0][-48:]})
    d[r[mn]] = float(r[mv].replace(',', ''))
tot = 0
for i, d in per.items():
    if 'at::' in d['name'] or 'pack' in d['name']: continue
    t = d.get('gpu__time_duration.sum', 0) / 1e3; tot += t
    print(f"{i:>4} {d['name']:48s} {t:8.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB  wr {d.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB")
print('total us', tot)
PY
for v in 3 5 6; do
  B200POSE_UPSAMPLE_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample_weight -c 8 --csv --log-file $O/upw_v$v.csv python tools/profile_step.py --passes 2 > $O/upw_v$v.log 2>&1
  echo "variant $v: $(grep upsample_weight $O/upw_v$v.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --cache-control none --clock-control none -k regex:conv_chain -c 10 --csv --log-file $O/chain_warm_traffic.csv python tools/profile_step.py --passes 2 > $O/chain_warm.log 2>&1
grep conv_chain $O/chain_warm_traffic.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | head -40
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $O/suite.txt 2>&1; tail -4 $O/suite.txt
ls $O
