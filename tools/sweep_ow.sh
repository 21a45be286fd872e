#!/bin/bash
# A/B sweep of the OW kernel scheduling parameters (RL_OW_SVC / RL_OW_LEAF / RL_OW_MINB) on C4 @50 spp and C5 (1920 wide) @64 spp
out=${1:-gpurun_out/sweep_ow.log}
: > $out
for wl in "C4 50" "C5 64"; do
  for minb in 3 4; do
    for sl in "12 12" "16 8" "16 12" "16 16" "20 12" "24 12" "20 16"; do
      set -- $sl
      echo -n "MINB=$minb " >> $out
      RL_OW_MINB=$minb RL_OW_SVC=$1 RL_OW_LEAF=$2 timeout 300 python tools/time_ow.py $wl >> $out 2>&1
    done
  done
done
cat $out
