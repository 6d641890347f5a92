#!/bin/bash
# Refresh of the step-kernel captures after the staged-target / occupancy changes, plus simple_spread (see gpu_profile_r02.sh).
mkdir -p gpurun_out/profiles_r02
P=gpurun_out/profiles_r02
: > $P/traffic_lines_b.json
cap() {
  name=$1; regex=$2; skip=$3; shift 3
  CS_PROFILE_BATCHES=64 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_$name "$@" > /dev/null 2>&1
  if [ -f /tmp/prof_$name.ncu-rep ]; then
    { echo "# ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 $*"; python tools/ncu_summary.py /tmp/prof_$name.ncu-rep; echo "# top source lines by stall samples"; python tools/ncu_lines.py /tmp/prof_$name.ncu-rep 25; echo "# top SASS instructions by stall samples"; python tools/ncu_sass.py /tmp/prof_$name.ncu-rep 16; } > $P/r02_${name}_ncu_full.txt 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[2])); u=dict(zip(rows[0],rows[1]))
sc={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
print(json.dumps({'$name': sum(float(d[k])*sc.get(u[k],1) for k in ('dram__bytes_read.sum','dram__bytes_write.sum'))}))" >> $P/traffic_lines_b.json
    rm -f /tmp/prof_$name.ncu-rep
  else
    echo "capture failed: $name" >> $P/errors.txt
  fi
}
cap c2 flight_tpe_group_kernel 12 python tools/profile_run.py c2 16
cap c2w flight_tpe_kernel 12 python tools/profile_run.py c2w 16
cap c3 flight_tpe_kernel 40 python tools/profile_run.py c3 50
cap spread spread_kernel 12 python tools/prof_spread.py
cap policy_tc policy_tc_kernel 12 python tools/prof_policy.py
bash tools/gpu_profile_r02_fix.sh > /dev/null 2>&1
head -8 $P/r02_bench_launches.csv; cat $P/traffic_lines_b.json; head -12 $P/r02_c2_ncu_full.txt
