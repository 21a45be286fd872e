#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_obj_ingest.py -q -p no:cacheprovider > $O/pytest_obj.log 2>&1; echo "obj rc=$?" >> $O/rc.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --deselect tests/test_gpu_obj_ingest.py > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -15 $O/pytest_obj.log; tail -4 $O/pytest_gpu.log
