#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 900 python tools/sweep_ow.py quick > $O/sweep.jsonl 2> $O/sweep.err; echo "sweep rc=$?" >> $O/rc.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
python tools/time_ow.py C4 100 ow.variant=5 > $O/plain_v5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ow_render -s 2 -c 1 -o $O/prof_v5_c4 python tools/time_ow.py C4 100 ow.variant=5 > $O/ncu_v5.log 2>&1
cat $O/rc.txt; tail -5 $O/pytest_gpu.log; cat $O/plain_v5.log
