#!/bin/bash
mkdir -p gpurun_out/r2
ncu --set full --clock-control none --import-source on -k regex:flight_map_tile_kernel -s 40 -c 1 -f -o /tmp/prof_c4i python tools/profile_run.py c4 50 > /dev/null 2>&1
python tools/ncu_lines.py /tmp/prof_c4i.ncu-rep 70 inst > gpurun_out/r2/c4_by_inst.txt
cat gpurun_out/r2/c4_by_inst.txt
