#!/bin/bash
# Round-2 GPU call 20: tensor-core stem of the encoder: parity tests (both stems), timing, launch list.
cd "$(dirname "$0")/.."
O=gpurun_out/r2t; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_dropin.py -m gpu -q 2>&1 | tail -4 | tee $O/tests_tc.txt
B200POSE_ENC_STEM=0 timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q 2>&1 | tail -2 | tee $O/tests_fp32.txt
for v in 1 0; do echo "enc_stem=$v: $(B200POSE_ENC_STEM=$v timeout 200 python tools/profile_encoder.py --time 2>&1 | grep 'ms per batch')" | tee -a $O/enc_stem.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/encoder_launches.csv python tools/profile_encoder.py --passes 1 > $O/encoder_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2t/encoder_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; kn = H.index('Kernel Name'); mv = H.index('Metric Value')
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= mv or 'at::' in r[kn] or 'pack' in r[kn]: continue
    n += 1
    if n <= 6: print(r[kn].split('(')[0][-40:], float(r[mv].replace(',', '')) / 1e3, 'us')
PY
