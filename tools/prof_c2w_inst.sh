#!/bin/bash
# per-source-line instruction counts of the 1M-env step kernel (sorted by executed instructions)
mkdir -p gpurun_out/r2
ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 12 -c 1 -f -o /tmp/prof_c2wi python tools/profile_run.py c2w 16 > /dev/null 2>&1
python tools/ncu_lines.py /tmp/prof_c2wi.ncu-rep 60 inst > gpurun_out/r2/c2w_by_inst.txt
cat gpurun_out/r2/c2w_by_inst.txt
