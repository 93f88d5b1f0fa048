#!/bin/bash
# Round-2 GPU call 1: baseline suite, flake hunt (no retry hook), compute-sanitizer logs, the never-run chain kernel, bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2a; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/suite_first.txt 2>&1
tail -3 $O/suite_first.txt
# 1. the experimental chained launch, guarded (a protocol bug traps instead of hanging; timeout on top)
B200POSE_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_refine.py -q -x -k chained > $O/chain_test.txt 2>&1
echo "chain test rc=$?" | tee -a $O/chain_test.txt; tail -5 $O/chain_test.txt
B200POSE_CONV_MODE=19 timeout 200 python tools/profile_step.py --passes 2 --time > $O/chain_time.txt 2>&1; tail -2 $O/chain_time.txt
timeout 200 python tools/profile_step.py --passes 2 --time > $O/default_time.txt 2>&1; tail -2 $O/default_time.txt
# 2. bench line
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json
# 3. flake hunt: the full GPU suite, consecutively, no retry hook
pass=0; fail=0
for i in $(seq 1 10); do
  if timeout 600 python -m pytest tests -m gpu -x -q > $O/suite_loop_$i.txt 2>&1; then pass=$((pass+1)); rm -f $O/suite_loop_$i.txt; else fail=$((fail+1)); fi
done
echo "full-suite loop: $pass passed, $fail failed" | tee $O/suite_loop_summary.txt
# the one test that failed once in round 1, with its predecessors, 300 times in one process, NaN-filled outputs
timeout 600 python tools/flake_hunt.py 300 > $O/ctx_loop.txt 2>&1; tail -2 $O/ctx_loop.txt
# 4. compute-sanitizer (small shapes): memcheck, racecheck, initcheck, synccheck
CS=/usr/local/cuda/bin/compute-sanitizer
SEL="tests/test_gpu_ops.py"
for tool in memcheck initcheck racecheck; do
  timeout 300 $CS --tool $tool --log-file $O/sanitizer_${tool}_ops.log --print-limit 50 python -m pytest $SEL -q -x -p no:cacheprovider > $O/sanitizer_${tool}_ops.out 2>&1
  echo "$tool ops rc=$?"; tail -2 $O/sanitizer_${tool}_ops.log
done
for tool in memcheck racecheck; do
  timeout 300 $CS --tool $tool --log-file $O/sanitizer_${tool}_refine.log --print-limit 50 python -m pytest tests/test_gpu_refine.py -q -x -p no:cacheprovider -k "refine_128x160_4x3 or host_entry or zero_iterations" > $O/sanitizer_${tool}_refine.out 2>&1
  echo "$tool refine rc=$?"; tail -2 $O/sanitizer_${tool}_refine.log
done
ls -la $O | head -40
