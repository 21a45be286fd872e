#!/bin/bash
# evidence B (one GPU): launch list of the bench command + ncu --set full of every kernel DESIGN.md talks about.
# Reports are summarised ON THE BOX (tools/ncu_summary.py, ncu_lines.py); only two .ncu-rep files travel back (64 MiB cap).
O=gpurun_out/$1; mkdir -p $O; T=/tmp/ncu_reps; mkdir -p $T
python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu > $O/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c4.csv python bench.py --steps 3 --warmup 3 --no-secondary --no-cpu > $O/ncu_launch.log 2>&1
cap() {  # tag, kernel regex, command...
  tag=$1; rx=$2; shift; shift
  "$@" > $O/plain_$tag.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o $T/$tag "$@" > $O/ncu_$tag.log 2>&1
  python tools/ncu_summary.py $T/$tag.ncu-rep $tag $O/ncu_$tag.json > /dev/null 2>> $O/summ.err
  python tools/ncu_lines.py $T/$tag.ncu-rep 40 $O/lines_$tag.json > $O/lines_$tag.txt 2>> $O/summ.err
}
cap ow_c4_500 k_ow_render python tools/time_ow.py C4 500
cap ow_c5_64 k_ow_render python tools/time_ow.py C5 64
# C5 at the bench configuration: a short metric list (ncu saves / restores the 9 GB partial buffer around every replay pass;
# --set full took 11 minutes here)
python tools/time_ow.py C5full 256 > $O/plain_ow_c5_4k_256.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:k_ow_render -s 2 -c 1 -o $T/ow_c5_4k_256 python tools/time_ow.py C5full 256 > $O/ncu_ow_c5_4k_256.log 2>&1
python tools/ncu_summary.py $T/ow_c5_4k_256.ncu-rep ow_c5_4k_256 $O/ncu_ow_c5_4k_256.json > /dev/null 2>> $O/summ.err
cap rtc_c3 k_rtc_render python tools/time_rtc.py C3
cap rtc_c2 k_rtc_render python tools/time_rtc.py C2
if [ "$2" = "variants" ]; then  # the pooled and the global wavefront variants (unchanged since evB2: only on request)
cap ow_pooled_c4_100 k_ow_render python tools/time_ow.py C4 100 ow.variant=6 ow.exit_min=24 ow.minb=3
cap ow_c4_100 k_ow_render python tools/time_ow.py C4 100
python tools/time_ow.py C4 20 ow.variant=7 > $O/plain_wf.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 400 -c 2 -o $T/wf_c4_20 python tools/time_ow.py C4 20 ow.variant=7 > $O/ncu_wf.log 2>&1
python tools/ncu_summary.py $T/wf_c4_20.ncu-rep wf_logic $O/ncu_wf_logic.json k_wf_logic > /dev/null 2>> $O/summ.err
python tools/ncu_summary.py $T/wf_c4_20.ncu-rep wf_trace $O/ncu_wf_trace.json k_wf_trace > /dev/null 2>> $O/summ.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_wf_ -c 2000 --csv --log-file $O/launches_wavefront_c4_20spp.csv python tools/time_ow.py C4 20 ow.variant=7 > $O/ncu_wf_launch.log 2>&1
fi
# the files bench.py reads for roofline.ncu / roofline.traffic
python tools/ncu_summary.py $T/ow_c4_500.ncu-rep C4 $O/ncu_C4.json > /dev/null 2>> $O/summ.err
python tools/ncu_summary.py $T/ow_c5_4k_256.ncu-rep C5 $O/ncu_C5.json > /dev/null 2>> $O/summ.err  # (short metric list)
python tools/ncu_summary.py $T/rtc_c3.ncu-rep C3 $O/ncu_C3.json > /dev/null 2>> $O/summ.err
python tools/ncu_summary.py $T/rtc_c2.ncu-rep C2 $O/ncu_C2.json > /dev/null 2>> $O/summ.err
cp $T/ow_c4_500.ncu-rep $T/ow_c5_64.ncu-rep $O/
ls -la $O $T; du -sh $O
