#!/bin/bash
# Round-2 GPU call 21: final state: full suite, smoke, bench (default, B=256, cfg4 B=64), launch lists (step + bench command),
# ncu --set full of the chained launch and of the other kernels of a step.
cd "$(dirname "$0")/.."
O=gpurun_out/r2u; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -5 $O/suite.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
timeout 600 python bench.py --config cfg1 --global-batch 256 --steps 5 --warmup 3 > $O/bench_cfg1_b256.json 2> $O/bench_cfg1_b256.err
timeout 600 python bench.py --config cfg4 --global-batch 64 --steps 5 --warmup 3 > $O/bench_cfg4_b64.json 2> $O/bench_cfg4_b64.err
python - <<'PY'
import json
for f in ('bench_default', 'bench_cfg1_b256', 'bench_cfg4_b64'):
    try:
        txt = open(f'gpurun_out/r2u/{f}.json').read(); d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1]); e = d.get('e2e') or {}
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(e.get('value', 0)), e.get('context'), 'frac', round(d['roofline']['frac'], 4), 'enc', (d.get('encoder') or {}).get('ms_per_batch'), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
    except Exception as ex: print(f, 'ERR', ex)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --fmaps hash --cpu-objects 1 --e2e-steps 1 > $O/ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file $O/launches_step.csv python tools/profile_step.py --passes 3 > $O/ncu_step.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 6 -c 1 -o $O/chain_ncu -f python tools/profile_step.py --passes 2 > $O/ncu_chain_full.log 2>&1
ncu -i $O/chain_ncu.ncu-rep --page raw --csv > $O/chain_ncu_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:"upsample_weight_kernel|lm_cluster|corr_lookup_win|corr_pool3|context_init|im2col_f1|flow_init|fmap_to_pxc|conv_umma_kernel" -s 12 -c 10 -o $O/misc_ncu -f python tools/profile_step.py --passes 2 > $O/ncu_misc_full.log 2>&1
ncu -i $O/misc_ncu.ncu-rep --page raw --csv > $O/misc_ncu_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:"enc_stem_im2col|in_stats1|in_apply|conv_umma2" -s 0 -c 8 -o $O/enc_ncu -f python tools/profile_encoder.py --passes 1 > $O/ncu_enc_full.log 2>&1
ncu -i $O/enc_ncu.ncu-rep --page raw --csv > $O/enc_ncu_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls $O
