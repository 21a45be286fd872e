#!/bin/bash
# A/B sweep of the OW kernel variants (RL_OW_KERNEL_V / RL_OW_MINB / RL_OW_SVC / RL_OW_LEAF / RL_OW_GENERIC) on C4 @50 spp and C5 (1920 wide) @8 spp
out=${1:-gpurun_out/sweep_ow.log}
: > $out
for wl in "C4 50" "C5 8"; do
  RL_OW_KERNEL_V=3 python tools/time_ow.py $wl >> $out 2>&1
  for gen in 0 1; do
  for minb in 3 4; do
    for sl in "16 12" "24 8" "16 32" "12 12" "16 16"; do
      set -- $sl
      echo -n "GEN=$gen MINB=$minb " >> $out
      RL_OW_GENERIC=$gen RL_OW_KERNEL_V=5 RL_OW_MINB=$minb RL_OW_SVC=$1 RL_OW_LEAF=$2 timeout 300 python tools/time_ow.py $wl >> $out 2>&1
    done
  done
  done
done
cat $out
