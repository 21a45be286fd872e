mkdir -p gpurun_out/s3
RL_OW_KERNEL_V=4 RL_OW_MINB=3 RL_OW_SVC=8 ncu --set full --clock-control none --import-source on -k regex:k_ow_render -c 1 -s 1 -o gpurun_out/s3/prof_v4_svc8 -f python tools/time_ow.py C4 20 > gpurun_out/s3/ncu_v4.log 2>&1
RL_OW_KERNEL_V=5 RL_OW_MINB=3 RL_OW_SVC=16 RL_OW_LEAF=8 ncu --set full --clock-control none --import-source on -k regex:k_ow_render -c 1 -s 1 -o gpurun_out/s3/prof_v5 -f python tools/time_ow.py C4 20 > gpurun_out/s3/ncu_v5.log 2>&1
ls -la gpurun_out/s3
