#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ow.py tests/test_gpu_ow_production.py tests/test_gpu_debug_build.py -q -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 300 python tools/time_ow.py C4 500 > $O/c4.log 2>&1
timeout 300 python tools/time_ow.py C5 64 > $O/c5.log 2>&1
cat $O/rc.txt; tail -4 $O/pytest.log; cat $O/c4.log $O/c5.log
