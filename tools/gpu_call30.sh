#!/bin/bash
# Round-2 GPU call 30: lookup modes 1 / 2 / 3 (late PDL trigger): step time, repeated; launch list of the neighbours.
cd "$(dirname "$0")/.."
O=gpurun_out/r3d; mkdir -p $O
export PYTHONDONTWRITEBYTECODE=1
for rep in 1 2; do for m in 1 2 3; do
  echo "lookup_mode=$m: $(B200POSE_LOOKUP_MODE=$m timeout 200 python tools/profile_step.py --passes 2 --time 2>&1 | grep 'ms per pass')" | tee -a $O/ab.txt
done; done
for m in 1 2; do
  B200POSE_LOOKUP_MODE=$m timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"corr_lookup|im2col_f1|conv_chain|upsample" -s 4 -c 12 --csv --log-file $O/nb_$m.csv python tools/profile_step.py --passes 1 > $O/nb_$m.log 2>&1
  echo "mode $m: $(grep -E 'corr_lookup|im2col|conv_chain|upsample' $O/nb_$m.csv | awk -F'","' '{print substr($5,1,18), $NF}' | tr -d '"' | tr '\n' ';')" | tee -a $O/ab.txt
done
