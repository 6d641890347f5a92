#!/bin/bash
mkdir -p gpurun_out/r2 gpurun_out/profiles_r02
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_full.log 2>&1; tail -2 gpurun_out/r2/pytest_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; tail -2 gpurun_out/r2/bench_n1.err
python bench.py --impl reference > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err
P=gpurun_out/profiles_r02
: > $P/traffic_lines_c.json
cap() {
  name=$1; regex=$2; skip=$3; shift 3
  CS_PROFILE_BATCHES=64 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_$name "$@" > /dev/null 2>&1
  if [ -f /tmp/prof_$name.ncu-rep ]; then
    { echo "# ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 $*"; python tools/ncu_summary.py /tmp/prof_$name.ncu-rep; echo "# top source lines by stall samples"; python tools/ncu_lines.py /tmp/prof_$name.ncu-rep 25; echo "# top source lines by executed instructions"; python tools/ncu_lines.py /tmp/prof_$name.ncu-rep 25 inst; echo "# top SASS instructions by stall samples"; python tools/ncu_sass.py /tmp/prof_$name.ncu-rep 12; } > $P/r02_${name}_ncu_full.txt 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[2])); u=dict(zip(rows[0],rows[1]))
sc={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
print(json.dumps({'$name': sum(float(d[k])*sc.get(u[k],1) for k in ('dram__bytes_read.sum','dram__bytes_write.sum'))}))" >> $P/traffic_lines_c.json
    rm -f /tmp/prof_$name.ncu-rep
  fi
}
cap c2 flight_tpe_group_kernel 12 python tools/profile_run.py c2 16
cap c2w flight_tpe_kernel 12 python tools/profile_run.py c2w 16
cap c3 flight_tpe_kernel 40 python tools/profile_run.py c3 50
cap c4 flight_map_tile_kernel 40 python tools/profile_run.py c4 50
bash tools/gpu_profile_r02_fix.sh > /dev/null 2>&1
cat $P/r02_bench_launches.csv | tail -3; cat $P/traffic_lines_c.json
python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_n1.json').read().strip().splitlines()[-1])
print('value %.4e frac %.3f us %.2f e2e %.4e' % (d['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['e2e']['value']))
for k,v in d['extra'].items():
    if isinstance(v,dict) and 'value' in v: print(k, '%.4e' % v['value'], 'frac %.3f' % v['roofline']['frac'], 'us %.2f' % v['us_per_launch'], 'e2e', v.get('e2e_value'))
    if isinstance(v,dict) and 'error' in v: print(k, v)
print([('%.3e' % r['agent_steps_per_s'], r['rows']) for r in d['extra']['policy']['runs']], d['extra']['policy']['device_loop']['us_per_step'])
print(d['cpu_baseline']['value'], d['extra']['c1']['reference_env']['value'], d['extra']['c1']['adapter_e1']['value'])
"
