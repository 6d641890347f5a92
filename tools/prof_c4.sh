#!/bin/bash
# c4: per-kernel launch list (warm caches) + one full capture of the map kernel and of the step kernel
mkdir -p gpurun_out
python tools/exp_c4.py 300 > gpurun_out/exp_c4.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --cache-control none -s 100 -c 8 --csv --log-file gpurun_out/c4_launches_warm.csv python tools/exp_c4.py 120 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_map_kernel -s 40 -c 1 -f -o gpurun_out/prof_c4_map python tools/exp_c4.py 60 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_kernel -s 40 -c 1 -f -o gpurun_out/prof_c4_step python tools/exp_c4.py 60 > /dev/null 2>&1
cat gpurun_out/exp_c4.log
