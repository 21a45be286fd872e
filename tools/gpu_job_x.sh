#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
for opts in "" "ow.steps=3" "ow.steps=3 ow.leaf_min=6" "ow.steps=3 ow.leaf_min=12" "" "ow.steps=3"; do
  timeout 120 python tools/time_ow.py C4 500 $opts >> $O/c4.log 2>&1
done
for opts in "" "ow.steps=3" "ow.steps=3 ow.leaf_min=12" ""; do
  timeout 120 python tools/time_ow.py C5 64 $opts >> $O/c5.log 2>&1
done
cat $O/c4.log $O/c5.log
