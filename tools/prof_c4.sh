#!/bin/bash
# ncu captures of the flight-variant kernels on the c4 workload: warm-cache full capture (what the timed loop sees) with
# top source lines, plus single-pass DRAM traffic (no replay, so no save/restore traffic).  Summaries only.
K=${1:-flight_map_tile_kernel}
mkdir -p gpurun_out/prof
ncu --set full --clock-control none --cache-control none --import-source on -k regex:$K -s 60 -c 1 -f -o gpurun_out/prof_c4_warm python tools/profile_run.py c4 70 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_c4_warm.ncu-rep > gpurun_out/prof/c4_warm_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_c4_warm.ncu-rep 70 >> gpurun_out/prof/c4_warm_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_c4_warm.ncu-rep 70 inst > gpurun_out/prof/c4_warm_byinst.txt 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,smsp__inst_executed.sum --clock-control none --cache-control none -k regex:"flight_map_tile_kernel|flight_tpe_kernel|flight_fused" -s 100 -c 8 --csv --log-file gpurun_out/prof/c4_traffic_warm.csv python tools/profile_run.py c4 70 > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/prof/c4_warm_summary.txt
