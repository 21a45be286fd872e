#!/bin/bash
O=gpurun_out/$1; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ow_production.py -q -p no:cacheprovider -k "schedule" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for opts in "" "ow.variant=7" "ow.variant=7 ow.slots=262144" "ow.variant=7 ow.slots=1048576" "ow.variant=7 ow.slots=2097152" "ow.variant=7 ow.exit_min=16"; do
  timeout 120 python tools/time_ow.py C4 500 $opts >> $O/c4.log 2>&1
done
for opts in "" "ow.variant=7" "ow.variant=7 ow.slots=1048576"; do
  timeout 120 python tools/time_ow.py C5 64 $opts >> $O/c5.log 2>&1
done
cat $O/rc.txt; tail -5 $O/pytest.log; cat $O/c4.log $O/c5.log
