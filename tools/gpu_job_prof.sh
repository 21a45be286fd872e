#!/bin/bash
# ncu --set full of the OW render kernel on the cover scene (100 spp), one launch: $1 = tag, rest = options
O=gpurun_out/prof; mkdir -p $O
tag=$1; shift
python tools/time_ow.py C4 100 "$@" > $O/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ow_render -s 2 -c 1 -o $O/prof_${tag}_c4 python tools/time_ow.py C4 100 "$@" > $O/ncu_$tag.log 2>&1
cat $O/plain_$tag.log; tail -2 $O/ncu_$tag.log
