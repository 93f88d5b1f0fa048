mkdir -p gpurun_out; L=gpurun_out/fg2.log
B200POSE_FG_UPSAMPLE=1 timeout 120 python -m pytest tests/test_gpu_refine.py -q -x -k "foreground or golden or batched or host_entry or zero_iter" 2>&1 | tail -2 > $L
for v in "B200POSE_FG_UPSAMPLE=0" "B200POSE_FG_UPSAMPLE=1 B200POSE_FG_BLOCKS=8" "B200POSE_FG_UPSAMPLE=1 B200POSE_FG_BLOCKS=6"; do echo "-- $v" >> $L; env $v timeout 100 python tools/profile_step.py --time --passes 2 2>&1 | grep "ms per pass" >> $L; done
cat $L
