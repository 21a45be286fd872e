#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 900 python tools/sweep_ow.py quick > $O/sweep.jsonl 2> $O/sweep.err; echo "sweep rc=$?" >> $O/rc.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -5 $O/pytest_gpu.log
