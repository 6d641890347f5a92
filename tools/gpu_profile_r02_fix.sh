#!/bin/bash
mkdir -p gpurun_out/profiles_r02
P=gpurun_out/profiles_r02
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flight_|policy_|search_" -s 64 -c 500 --csv --log-file $P/r02_bench_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1
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
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:flight_|policy_|search_ -s 64 -c 500 python bench.py --steps 2 --warmup 3 --no-extra\n# (our kernels after the 64 reset launches: warm-up, the timed K-step graph replays, then the e2e forms; cold-cache, serialised)\nkernel,launches,total_us,mean_us\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write('"%s",%d,%.2f,%.2f\n' % (k, n, t, t / n))
PY
rm -f $P/r02_bench_launches_raw.csv
cat $P/r02_bench_launches.csv
