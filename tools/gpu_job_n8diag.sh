#!/bin/bash
N=$1; O=gpurun_out/$2; mkdir -p $O
for mode in device; do
RL_MGPU=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary > $O/bench_$mode.json 2> $O/bench_$mode.err; echo "$mode rc=$?" >> $O/rc.txt
done
cat $O/rc.txt
