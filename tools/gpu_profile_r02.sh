#!/bin/bash
# Round-2 ncu evidence behind profiles/r02_* (run under gpurun, 1 GPU): launch lists + one full capture per hot kernel.
# The reports are summarised on the GPU box (they exceed the copy-back limit); only text comes back.
mkdir -p gpurun_out/profiles_r02
P=gpurun_out/profiles_r02
# launch list of the default bench command (kernel shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r02_bench_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/profiles_r02/r02_bench_launches_raw.csv")) if len(r) > 5]
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            k = d["Kernel Name"][:100]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
with open("gpurun_out/profiles_r02/r02_bench_launches.csv", "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-extra (cold-cache, serialised)\nkernel,launches,total_us,mean_us\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write('"%s",%d,%.2f,%.2f\n' % (k, n, t, t / n))
PY
rm -f $P/r02_bench_launches_raw.csv
cap() {  # name, kernel regex, skip, target command...
  name=$1; regex=$2; skip=$3; shift 3
  CS_PROFILE_BATCHES=64 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_$name "$@" > /dev/null 2>&1
  if [ -f /tmp/prof_$name.ncu-rep ]; then
    { echo "# ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 $*"; python tools/ncu_summary.py /tmp/prof_$name.ncu-rep; echo "# top source lines by stall samples"; python tools/ncu_lines.py /tmp/prof_$name.ncu-rep 25; echo "# top SASS instructions by stall samples"; python tools/ncu_sass.py /tmp/prof_$name.ncu-rep 16; } > $P/r02_${name}_ncu_full.txt 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[2])); u=dict(zip(rows[0],rows[1]))
sc={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
print(json.dumps({'$name': sum(float(d[k])*sc.get(u[k],1) for k in ('dram__bytes_read.sum','dram__bytes_write.sum'))}))" >> $P/traffic_lines.json
    rm -f /tmp/prof_$name.ncu-rep
  else
    echo "capture failed: $name" >> $P/errors.txt
  fi
}
cap c2 flight_tpe_group_kernel 12 python tools/profile_run.py c2 16
cap c2w flight_tpe_kernel 12 python tools/profile_run.py c2w 16
cap c3 flight_tpe_kernel 40 python tools/profile_run.py c3 50
cap c4 flight_map_tile_kernel 40 python tools/profile_run.py c4 50
cap c4_step flight_tpe_kernel 40 python tools/profile_run.py c4 50
cap c5 search_kernel 20 python tools/profile_run.py c5 30
cap policy_tc policy_tc_kernel 3 python tools/prof_policy.py
cap pack_pool flight_pack_pool_kernel 3 python tools/probes/e2e_trace.py
ls -la $P
