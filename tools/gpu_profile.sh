#!/bin/bash
# ncu captures behind profiles/: launch list + one full capture per hot kernel (run under gpurun, 1 GPU).
# c2 uses all 64 batches so that, like in bench.py, every launch reads its state from HBM.
mkdir -p gpurun_out
CS_PROFILE_BATCHES=64 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:flight_tpe_group_kernel -s 10 -c 30 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_run.py c2 45 > /dev/null 2>&1
CS_PROFILE_BATCHES=64 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:flight_tpe_kernel -s 200 -c 64 --csv --log-file gpurun_out/launches_c2s.csv python tools/profile_run.py c2s 6 > /dev/null 2>&1
for wl in c3 c4 c5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/launches_$wl.csv python tools/profile_run.py $wl 30 > /dev/null 2>&1
done
CS_PROFILE_BATCHES=64 ncu --set full --clock-control none --import-source on -k regex:flight_tpe_group_kernel -s 12 -c 1 -f -o gpurun_out/prof_c2 python tools/profile_run.py c2 16 > /dev/null 2>&1
CS_PROFILE_BATCHES=64 ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 200 -c 1 -f -o gpurun_out/prof_c2s python tools/profile_run.py c2s 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 40 -c 1 -f -o gpurun_out/prof_c3 python tools/profile_run.py c3 20 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_map_kernel -s 40 -c 1 -f -o gpurun_out/prof_c4 python tools/profile_run.py c4 50 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 40 -c 1 -f -o gpurun_out/prof_c4_step python tools/profile_run.py c4 50 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 12 -c 1 -f -o gpurun_out/prof_c2w python tools/profile_run.py c2w 16 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 20 -c 1 -f -o gpurun_out/prof_c5 python tools/profile_run.py c5 30 > /dev/null 2>&1
ls gpurun_out | tail -12
# the reports are too large to travel back: keep the summaries only
python tools/make_profile_summary.py r01 gpurun_out/profiles_r01 > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
