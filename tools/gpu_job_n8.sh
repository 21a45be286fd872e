#!/bin/bash
N=$1; O=gpurun_out/$2; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 8 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "benchN rc=$?" >> $O/rc.txt
if [ "$3" = "multi" ]; then timeout 300 python tools/time_multi.py $N > $O/time_multi.log 2>&1; echo "multi rc=$?" >> $O/rc.txt; cat $O/time_multi.log; fi
cat $O/rc.txt; tail -3 $O/bench_n$N.err
