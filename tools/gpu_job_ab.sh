#!/bin/bash
# A/B of experiment builds (tools/build_alt.py): tools/gpu_job_ab.sh OUT lib1 lib2 ...   ("-" = the product library)
O=gpurun_out/$1; mkdir -p $O; shift
for rep in 1 2; do
for l in "$@"; do
  a=""; [ "$l" != "-" ] && a="lib=$l"
  for w in "C4 500" "C5 64"; do
    echo -n "[$l] " >> $O/ab.log; timeout 300 python tools/time_ow.py $w $a >> $O/ab.log 2>&1
  done
done
done
cat $O/ab.log
