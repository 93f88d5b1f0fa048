#!/bin/bash
# Round-2 GPU call 10 (2 GPUs): the torchrun path of bench.py (both arms) as the driver launches it.
cd "$(dirname "$0")/.."
O=gpurun_out/r2j; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi topo -m > $O/topo.txt 2>&1
ls /sys/devices/system/node/ > $O/numa_nodes.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; cut -c1-400 $O/bench_n2.json; tail -3 $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; cut -c1-200 $O/bench_ref_n2.json
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; cut -c1-200 $O/bench_n1.json
python -c "
import json
for f in ('bench_n1','bench_n2'):
    d=json.load(open('$O/'+f+'.json')); print(f, d['value'], d['e2e']['value'], d['e2e']['numa'], d['e2e']['h2d_gbs_this_rank'])"
ls $O
