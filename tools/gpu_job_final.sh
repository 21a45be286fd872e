# final evidence run on ONE GPU: tests, benches, launch list, ncu --set full at the bench configuration
set -x
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_c4.json 2> $O/bench_ref_c4.err; echo "ref rc=$?"
for wl in C1 C2 C3 C5; do python bench.py --workload $wl --steps 3 > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err; echo "$wl rc=$?"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $O/ncu_launches.log 2>&1
python tools/time_ow.py C4 500 > $O/plain_c4_500.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ow_render -c 1 -s 1 -o $O/prof_ow_c4_500 -f python tools/time_ow.py C4 500 > $O/ncu_c4_500.log 2>&1
python tools/time_e2e.py C5 > $O/plain_c5_full.log 2>&1 && ncu --set full --clock-control none -k regex:k_ow_render -c 1 -s 1 -o $O/prof_ow_c5_full -f python tools/time_e2e.py C5 > $O/ncu_c5_full.log 2>&1
python tools/time_rtc.py C3 > $O/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_rtc_render -c 1 -s 1 -o $O/prof_rtc_c3 -f python tools/time_rtc.py C3 > $O/ncu_c3.log 2>&1
ncu --set full --clock-control none -k regex:"k_morton|k_radix_sort|k_hierarchy|k_refit|k_pack|k_centroid" -c 6 -o $O/prof_lbvh_c3 -f python tools/time_rtc.py C3 > $O/ncu_lbvh.log 2>&1
cat $O/plain_*.log; tail -3 $O/pytest_gpu.log
