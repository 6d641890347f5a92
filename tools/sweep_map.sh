#!/bin/bash
# map kernel A/B: TMA-tile vs direct sweep, timed on c4; warm-cache launch list and one full capture
mkdir -p gpurun_out
{
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== tma"; CS_MAP_TMA=1 python tools/exp_c4.py 300
echo "== direct"; python tools/exp_c4.py 300
} > gpurun_out/sweep_map.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --cache-control none -s 100 -c 4 --csv --log-file gpurun_out/c4_launches_warm.csv python tools/exp_c4.py 120 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:flight_map -s 40 -c 1 -f -o gpurun_out/prof_c4_map python tools/exp_c4.py 60 > /dev/null 2>&1
cat gpurun_out/sweep_map.log
