#!/bin/bash
# Round-2 GPU call 8: full suite incl. LM backward, final ncu evidence (launch list + --set full of the chained launch and the
# other kernels), bench lines for cfg1 / cfg3 / cfg1 B=256.
cd "$(dirname "$0")/.."
O=gpurun_out/r2h; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/suite.txt 2>&1; tail -8 $O/suite.txt
timeout 200 python tools/profile_step.py --passes 2 --time > $O/time_default.txt 2>&1; grep "ms per pass" $O/time_default.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; cut -c1-400 $O/bench_cfg1.json; tail -2 $O/bench_cfg1.err
# the same command under ncu: launch list (cold cache, serialised: compare SHARES)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --fmaps hash --cpu-objects 1 --e2e-steps 1 > $O/ncu_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file $O/launches_step.csv python tools/profile_step.py --passes 3 > $O/ncu_step.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 6 -c 1 -o $O/chain_ncu -f python tools/profile_step.py --passes 2 > $O/ncu_chain_full.log 2>&1
ncu -i $O/chain_ncu.ncu-rep --page raw --csv > $O/chain_ncu_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:"upsample_weight_kernel|lm_multi|corr_lookup_win|corr_pool3|context_init|im2col|flow_init|fmap_to_pxc|fg_fill|conv_umma_kernel" -s 12 -c 10 -o $O/misc_ncu -f python tools/profile_step.py --passes 2 > $O/ncu_misc_full.log 2>&1
ncu -i $O/misc_ncu.ncu-rep --page raw --csv > $O/misc_ncu_raw.csv 2>/dev/null
timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; cut -c1-300 $O/bench_cfg3.json
timeout 600 python bench.py --config cfg1 --global-batch 256 --steps 5 --warmup 3 > $O/bench_cfg1_b256.json 2> $O/bench_cfg1_b256.err; cut -c1-300 $O/bench_cfg1_b256.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
rm -f $O/*.ncu-rep
ls $O
