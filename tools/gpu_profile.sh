#!/bin/bash
# ncu captures used for profiles/: launch list + one full capture per hot kernel (run under gpurun, 1 GPU)
mkdir -p gpurun_out
for wl in c2 c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_$wl.csv python tools/profile_run.py $wl 30 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:flight_kernel -s 40 -c 1 -f -o gpurun_out/prof_c2 python tools/profile_run.py c2 20 > /dev/null 2>&1
CS_BENCH_LPE=2 ncu --set full --clock-control none --import-source on -k regex:flight_kernel -s 40 -c 1 -f -o gpurun_out/prof_c3 python tools/profile_run.py c3 20 > /dev/null 2>&1
CS_BENCH_LPE=16 ncu --set full --clock-control none --import-source on -k regex:flight_kernel -s 40 -c 1 -f -o gpurun_out/prof_c4 python tools/profile_run.py c4 50 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 20 -c 1 -f -o gpurun_out/prof_c5 python tools/profile_run.py c5 30 > /dev/null 2>&1
ls -la gpurun_out | tail -12
