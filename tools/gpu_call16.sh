#!/bin/bash
# Round-2 GPU call 16: two half batches on two streams vs one batch; bench default (NVML clock sampler).
cd "$(dirname "$0")/.."
O=gpurun_out/r2p; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python tools/split_streams.py --parts 4 > $O/split_skew1.txt 2>&1; cat $O/split_skew1.txt
timeout 300 python tools/split_streams.py --parts 4 --skew 0 > $O/split_skew0.txt 2>&1; cat $O/split_skew0.txt
timeout 300 python tools/split_streams.py --batch 64 --parts 4 > $O/split_b64.txt 2>&1; cat $O/split_b64.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python -c "
import json; d=json.load(open('$O/bench_default.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])"
