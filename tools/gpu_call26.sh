#!/bin/bash
# Round-2 GPU call 26 (8 GPUs): bench.py under torchrun as the driver launches it (configs[2]: 256 objects over 8 GPUs).
cd "$(dirname "$0")/.."
O=gpurun_out/r2z; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc > $O/nproc.txt; cat $O/nproc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; tail -3 $O/bench_n8.err
python - <<'PY'
import json
txt = open('gpurun_out/r2z/bench_n8.json').read()
d = json.loads([l for l in txt.splitlines() if l.startswith('{')][-1]); e = d['e2e']
print('N=8 value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(e['value']), e['context'], 'cpus', e['host_cpus'], e['numa'])
for r in e['variants']: print('   ', r['context'], round(r['value']), round(r['ms_per_step'], 2), round(r['h2d_gbs_this_rank'], 1))
print(d['clocks'])
PY
