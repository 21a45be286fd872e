set -x
mkdir -p gpurun_out/s2
python -m pytest tests -m gpu -x -q > gpurun_out/s2/pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py --workload C5 --steps 2 --warmup 3 > gpurun_out/s2/bench_c5_n1.json 2> gpurun_out/s2/bench_c5_n1.err; echo "c5 rc=$?"
python bench.py --workload C3 --steps 5 --warmup 3 > gpurun_out/s2/bench_c3_n1.json 2> gpurun_out/s2/bench_c3_n1.err; echo "c3 rc=$?"
python bench.py --workload C2 --steps 5 --warmup 3 > gpurun_out/s2/bench_c2_n1.json 2> gpurun_out/s2/bench_c2_n1.err; echo "c2 rc=$?"
python bench.py --workload C1 --steps 5 --warmup 3 > gpurun_out/s2/bench_c1_n1.json 2> gpurun_out/s2/bench_c1_n1.err; echo "c1 rc=$?"
python tools/time_rtc.py C3 > gpurun_out/s2/plain_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_rtc_render -c 1 -s 1 -o gpurun_out/s2/prof_rtc_c3 -f python tools/time_rtc.py C3 > gpurun_out/s2/ncu_c3.log 2>&1
python tools/time_rtc.py C2 > gpurun_out/s2/plain_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_rtc_render -c 1 -s 1 -o gpurun_out/s2/prof_rtc_c2 -f python tools/time_rtc.py C2 > gpurun_out/s2/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s2/launches_rtc_c3.csv python tools/time_rtc.py C3 > gpurun_out/s2/ncu_l_c3.log 2>&1
python tools/time_ow.py C5 8 > gpurun_out/s2/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ow_render -c 1 -s 1 -o gpurun_out/s2/prof_ow_c5 -f python tools/time_ow.py C5 8 > gpurun_out/s2/ncu_c5.log 2>&1
ls -la gpurun_out/s2
