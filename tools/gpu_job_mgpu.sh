#!/bin/bash
# $1 = number of GPUs, $2 = output tag
N=$1; O=gpurun_out/$2; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 600 python -m pytest tests/test_gpu_ow_production.py -q -k "multi_gpu or identical or schedule" -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python tools/time_multi.py $N > $O/time_multi.log 2>&1; echo "multi rc=$?" >> $O/rc.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$?" >> $O/rc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "benchN rc=$?" >> $O/rc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 0 --impl reference --cpu-seconds 3 > $O/ref_n$N.json 2> $O/ref_n$N.err; echo "refN rc=$?" >> $O/rc.txt
cat $O/rc.txt $O/time_multi.log; tail -3 $O/pytest.log; tail -3 $O/bench_n$N.err
