#!/bin/bash
# c4 experiments: persisting-L2 window over the probability maps
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
{
python tools/exp_c4.py 300
for cfg in "32 0.3" "48 0.5" "64 0.6" "64 0.8" "80 0.7" "80 1.0" "96 0.8" "120 1.0"; do
  set -- $cfg
  echo "== persist $1 MB hit $2"
  CS_L2_PERSIST_MB=$1 CS_L2_HIT=$2 python tools/exp_c4.py 300
done
} > gpurun_out/exp_c4.log 2>&1
CS_L2_PERSIST_MB=64 CS_L2_HIT=0.6 ncu --metrics $M --clock-control none --cache-control none -k regex:flight_kernel -s 100 -c 4 --csv --log-file gpurun_out/exp_c4_warm_p64.csv python tools/exp_c4.py 120 > /dev/null 2>&1
CS_L2_PERSIST_MB=80 CS_L2_HIT=1.0 ncu --metrics $M --clock-control none --cache-control none -k regex:flight_kernel -s 100 -c 4 --csv --log-file gpurun_out/exp_c4_warm_p80.csv python tools/exp_c4.py 120 > /dev/null 2>&1
cat gpurun_out/exp_c4.log
