mkdir -p gpurun_out/s5
python tools/time_ow.py C5 16 > gpurun_out/s5/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ow_render -c 1 -s 1 -o gpurun_out/s5/prof_ow_c5 -f python tools/time_ow.py C5 16 > gpurun_out/s5/ncu_c5.log 2>&1
cat gpurun_out/s5/plain_c5.log
