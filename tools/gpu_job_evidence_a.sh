#!/bin/bash
# evidence A (one GPU, small outputs): smoke, the GPU test suite, bench lines for every BASELINE config + the reference arm
O=gpurun_out/$1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_C4_n1.json 2> $O/bench_C4_n1.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_reference_C4.json 2> $O/bench_reference_C4.err
for wl in C1 C2 C3 C5; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err; echo "bench $wl rc=$?" >> $O/rc.txt
done
timeout 300 python tools/time_fixed_cost.py > $O/fixed_cost.log 2>&1
cat $O/rc.txt; tail -3 $O/pytest_gpu.log
