#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for rep in 1 2; do for l in - csg1 csg6; do a=""; [ "$l" != "-" ] && a="lib=$l"; echo -n "[$l] " >> $O/csg.log; timeout 120 python tools/time_rtc.py CSG $a >> $O/csg.log 2>&1; done; done
tail -5 $O/pytest_gpu.log; cat $O/csg.log $O/rc.txt
