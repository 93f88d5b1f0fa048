mkdir -p gpurun_out; L=gpurun_out/fg3.log
timeout 100 python -m pytest tests/test_gpu_ops.py tests/test_gpu_refine.py -q -x -k "upsample or weight or foreground or golden or batched or host_entry" 2>&1 | tail -2 > $L
B200POSE_FG_UPSAMPLE=1 timeout 60 python -m pytest tests/test_gpu_refine.py -q -x -k "foreground or batched" 2>&1 | tail -1 >> $L
for v in "B200POSE_FG_UPSAMPLE=0" "B200POSE_FG_UPSAMPLE=1 B200POSE_FG_BLOCKS=8"; do echo "-- $v" >> $L; env $v timeout 100 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; done
cat $L
