#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_rtc.py -q -p no:cacheprovider -x > $O/pytest_rtc.log 2>&1; echo "rtc rc=$?" >> $O/rc.txt
timeout 300 python -m pytest tests/test_gpu_debug_build.py -q -p no:cacheprovider > $O/pytest_dbg.log 2>&1; echo "dbg rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -12 $O/pytest_rtc.log; tail -12 $O/pytest_dbg.log
