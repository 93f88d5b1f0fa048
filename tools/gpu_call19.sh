#!/bin/bash
# Round-2 GPU call 19: encoder in chunks of pairs (L2-resident normalisation passes).
cd "$(dirname "$0")/.."
O=gpurun_out/r2s; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
for c in 0 16 8 4 2; do
  echo "enc_chunk=$c: $(B200POSE_ENC_CHUNK=$c timeout 200 python tools/profile_encoder.py --time 2>&1 | grep 'ms per batch')" | tee -a $O/enc_chunk.txt
done
B200POSE_ENC_CHUNK=8 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/encoder_launches_chunk8.csv python tools/profile_encoder.py --passes 1 > $O/encoder_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2s/encoder_launches_chunk8.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; kn = H.index('Kernel Name'); mn = H.index('Metric Name'); mv = H.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv or 'at::' in r[kn] or 'pack' in r[kn]: continue
    a = agg.setdefault(r[kn].split('(')[0][-40:], collections.Counter()); a[r[mn]] += float(r[mv].replace(',', ''))
for k, v in agg.items(): print(f"{k:42s} {v['gpu__time_duration.sum'] / 1e3:8.1f} us  rd {v['dram__bytes_read.sum'] / 1e6:8.0f} MB  wr {v['dram__bytes_write.sum'] / 1e6:8.0f} MB")
PY
B200POSE_ENC_CHUNK=8 timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q 2>&1 | tail -2
