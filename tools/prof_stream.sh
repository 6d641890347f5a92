#!/bin/bash
# one full ncu capture of the streaming step kernel on a bench workload ($1, default c2w); summary only
wl=${1:-c2w}
mkdir -p gpurun_out/r2
out=gpurun_out/r2/prof_stream_$wl.txt
CS_PROFILE_BATCHES=64 ncu --set full --clock-control none --import-source on -k regex:flight_stream_kernel -s 12 -c 1 -f -o /tmp/prof_stream python tools/profile_run.py $wl 16 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_stream.ncu-rep > $out
python tools/ncu_lines.py /tmp/prof_stream.ncu-rep 40 >> $out
python tools/ncu_sass.py /tmp/prof_stream.ncu-rep 40 >> $out
cat $out
