#!/bin/bash
# Round-2 GPU call 17 (4 GPUs): bench.py under torchrun as the driver launches it; host-gather e2e at 4 ranks.
cd "$(dirname "$0")/.."
O=gpurun_out/r2q; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc > $O/nproc.txt; cat $O/nproc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; tail -3 $O/bench_n4.err
python -c "
import json; d=json.load(open('$O/bench_n4.json')); e=d['e2e']; print('N=4 value', d['value'], 'ms', d['ms_per_step']); print('e2e', e['value'], e['context'], 'cpus', e['host_cpus'], e['numa']); [print('  ', r) for r in e['variants']]; print(d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --config cfg3 --steps 10 --warmup 3 > $O/bench_cfg3_n4.json 2> $O/bench_cfg3_n4.err
python -c "
import json; d=json.load(open('$O/bench_cfg3_n4.json')); e=d['e2e']; print('cfg3 N=4 value', d['value'], 'e2e', e['value'], e['context'])"
