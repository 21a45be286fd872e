#!/bin/bash
# GPU test suite + the OW timings (C4 500 spp, C5 1920 x 64 spp, C5 at the bench configuration)
O=gpurun_out/$1; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
for w in "C4 500" "C5 64" "C5full 256"; do timeout 300 python tools/time_ow.py $w >> $O/time.log 2>&1; done
tail -4 $O/pytest_gpu.log; cat $O/time.log $O/rc.txt
