#!/bin/bash
# ncu captures of the fused flight kernel on the c4 workload: warm-cache full capture (what the timed loop sees) with
# top source lines, plus the cold-cache DRAM traffic of one launch.  Summaries only (the reports exceed the copy-back limit).
mkdir -p gpurun_out/prof
ncu --set full --clock-control none --cache-control none --import-source on -k regex:flight_fused_kernel -s 60 -c 1 -f -o gpurun_out/prof_c4_warm python tools/profile_run.py c4 70 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_c4_warm.ncu-rep > gpurun_out/prof/c4_warm_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_c4_warm.ncu-rep 45 >> gpurun_out/prof/c4_warm_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_c4_warm.ncu-rep 30 inst > gpurun_out/prof/c4_warm_byinst.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_fused_kernel -s 60 -c 1 -f -o gpurun_out/prof_c4_cold python tools/profile_run.py c4 70 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_c4_cold.ncu-rep > gpurun_out/prof/c4_cold_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/prof/c4_warm_summary.txt
