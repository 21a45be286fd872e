#!/bin/bash
O=gpurun_out/$1; mkdir -p $O; shift
for l in "$@"; do
  a=""; [ "$l" != "-" ] && a="lib=$l"
  echo "=== $l" >> $O/tail.log; timeout 300 python tools/time_fixed_cost.py $a $QUICK >> $O/tail.log 2>&1
done
cat $O/tail.log
