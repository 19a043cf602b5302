#!/bin/bash
# One GPU call's worth of evidence for profiles/: ncu --set full captures of one steady-state source frame (pack, search,
# two warps) at three resolutions, the launch list of the bench command, the search generations side by side.
# Every profiled command first runs once without ncu. Usage (on the GPU box): bash tools/prof_round.sh
set -u
mkdir -p gpurun_out
for cfg in "1920 1080 0 1080p-nv12" "3840 2160 1 4k-p010" "7680 4320 1 8k-p010"; do
  set -- $cfg
  HR_SEARCH_GEN=3 python tools/prof_case.py $1 $2 $3 5 2 4 > /dev/null || exit 1
  HR_SEARCH_GEN=3 ncu --set full --clock-control none --import-source on --launch-skip 9 -c 4 -f -o gpurun_out/r02_$4 python tools/prof_case.py $1 $2 $3 5 2 4 > gpurun_out/ncu_$4.log 2>&1
  tail -1 gpurun_out/ncu_$4.log
  python profiles/extract.py gpurun_out/r02_$4.ncu-rep > gpurun_out/r02_ncu_$4.txt
  if [ "$4" = "1080p-nv12" ]; then
    # instructions executed and stall samples per source line of the search kernel, stall reasons of the launch
    ncu -i gpurun_out/r02_$4.ncu-rep --page source --csv --kernel-name regex:flow_search > gpurun_out/src.csv 2>/dev/null
    (cd gpurun_out && cuobjdump -xelf all ../mpv-frame-interpolator_b200/csrc/libhopperrender_cuda.so > /dev/null && nvdisasm -g -c hr_cuda.sm_100a.cubin > all.sass 2>/dev/null)
    python profiles/sass_by_line.py gpurun_out/src.csv gpurun_out/all.sass _Z19flow_search3_kernelILi5ELb0EEv10FlowParams 40 > gpurun_out/r02_search3_source_counters.txt 2>&1
    ncu -i gpurun_out/r02_$4.ncu-rep --page raw --csv --kernel-name regex:flow_search 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,r=rows[0],rows[2]; d=dict(zip(h,r))
for k in h:
    if ('issue_stalled' in k and 'per_issue_active' in k) or k in ('smsp__average_warp_latency_per_inst_issued.ratio','smsp__inst_executed.sum','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct'): print('%-90s %s' % (k, d[k]))
" >> gpurun_out/r02_search3_source_counters.txt
    rm -f gpurun_out/src.csv gpurun_out/all.sass gpurun_out/*.cubin
  fi
  rm -f gpurun_out/r02_$4.ncu-rep
done
python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_steps6.json 2> gpurun_out/bench_steps6.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_steps6.csv python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for cfg in "1920 1080 0 5 8 16" "1280 720 0 5" "3840 2160 1 5"; do
  python tools/diag_search2.py $cfg
done > gpurun_out/search_generations.txt 2>&1
echo "== generation 1, 1080p NV12 R = 5, timeline (DBG instantiation, tools/diag_timeline.py)" >> gpurun_out/search_generations.txt
HR_SEARCH_GEN=1 python tools/diag_timeline.py 2>&1 | head -40 >> gpurun_out/search_generations.txt
for g in 2 3; do echo "== generation $g, 1080p NV12 R = 5, step timeline (DBG instantiation)"; GEN=$g STAGED=0 python tools/diag_search2_where.py 1920 1080 0 5; done >> gpurun_out/search_generations.txt 2>&1
for g in 1 3; do for r in 5 16; do echo "== pipelined device loop, generation $g R = $r"; HR_SEARCH_GEN=$g python tools/diag_pipeline.py 1920 1080 0 $r 1000 | tail -3; done; done >> gpurun_out/search_generations.txt 2>&1
