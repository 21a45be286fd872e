#!/bin/bash
# final evidence on ONE GPU: full GPU test suite, smoke, bench lines (C4 headline + C1/C2/C3/C5), launch list, ncu --set full
O=gpurun_out/$1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_C4_n1.json 2> $O/bench_C4_n1.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_reference_C4.json 2> $O/bench_reference_C4.err
for wl in C1 C2 C3 C5; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err; echo "bench $wl rc=$?" >> $O/rc.txt
done
# launch list of the bench command (cold-cache, serialised: compare SHARES)
python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu > $O/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c4.csv python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu > $O/ncu_launch.log 2>&1
# ncu --set full of the dominant kernel at the bench configuration (C4 500 spp), C5 (1920 wide, 64 spp) and C3
python tools/time_ow.py C4 500 > $O/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ow_render -s 2 -c 1 -o $O/prof_ow_c4_500 python tools/time_ow.py C4 500 > $O/ncu_c4.log 2>&1
python tools/time_ow.py C5 64 > $O/plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ow_render -s 2 -c 1 -o $O/prof_ow_c5 python tools/time_ow.py C5 64 > $O/ncu_c5.log 2>&1
python tools/time_rtc.py C3 > $O/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rtc_render -s 2 -c 1 -o $O/prof_rtc_c3 python tools/time_rtc.py C3 > $O/ncu_c3.log 2>&1
python tools/time_ow.py C4 100 ow.variant=6 ow.exit_min=24 ow.minb=3 > $O/plain_v6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ow_render -s 2 -c 1 -o $O/prof_ow_pooled_c4_100 python tools/time_ow.py C4 100 ow.variant=6 ow.exit_min=24 ow.minb=3 > $O/ncu_v6.log 2>&1
python tools/time_ow.py C4 100 > $O/plain_c4_100.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_ow_render -s 2 -c 1 -o $O/prof_ow_c4_100 python tools/time_ow.py C4 100 > $O/ncu_c4_100.log 2>&1
cat $O/rc.txt; tail -3 $O/pytest_gpu.log
