#!/bin/bash
# round 2, first GPU call: does v6 work at all -> sweep -> full GPU test suite -> bench lines -> ncu
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 900 python tools/sweep_ow.py quick > $O/sweep.jsonl 2> $O/sweep.err; echo "sweep rc=$?" >> $O/rc.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -5 $O/pytest_gpu.log
