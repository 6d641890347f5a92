#!/bin/bash
# one full ncu capture of the thread-per-env step kernel on the 1M-env workload (c2w) and on c4's 16384 envs
mkdir -p gpurun_out
CS_TPE_K=${CS_TPE_K:-1} ncu --set full --clock-control none --import-source on -k regex:flight_tpe_kernel -s 12 -c 1 -f -o gpurun_out/prof_c2w_tpe python tools/profile_run.py c2w 16 > gpurun_out/prof_step.log 2>&1
for k in 1 2 4; do CS_TPE_K=$k python tools/exp_c4.py 200 >> gpurun_out/prof_step.log 2>&1; done
tail -4 gpurun_out/prof_step.log
