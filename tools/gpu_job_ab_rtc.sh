#!/bin/bash
# A/B of experiment builds on the RTC workloads: tools/gpu_job_ab_rtc.sh OUT lib1 lib2 ...   ("-" = the product library)
O=gpurun_out/$1; mkdir -p $O; shift
for rep in 1 2; do
for l in "$@"; do
  a=""; [ "$l" != "-" ] && a="lib=$l"
  for w in C1 C2 C3; do
    echo -n "[$l] " >> $O/ab.log; timeout 300 python tools/time_rtc.py $w $a >> $O/ab.log 2>&1
  done
done
done
cat $O/ab.log
