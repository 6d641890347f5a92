"""Turns the ncu reports under gpurun_out/ into the small committed summaries under profiles/.
usage: python tools/make_profile_summary.py r01 [output dir, default profiles/]
(on the GPU box: write into gpurun_out/profiles_<tag>/ and delete the .ncu-rep files, which exceed the copy-back limit)"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
traffic = {}
for wl in ("c2", "c2s", "c2w", "c3", "c4", "c4_step", "c5"):
    rep = os.path.join(src, "prof_%s.ncu-rep" % wl)
    if os.path.exists(rep):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
        with open(os.path.join(dst, "%s_%s_ncu_full.txt" % (tag, wl)), "w") as fh:
            fh.write("# ncu --set full --clock-control none --import-source on, one launch of the %s workload\n" % wl)
            fh.write(txt + "\n# top source lines by stall samples\n" + lines)
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[k]) * scale.get(u[k], 1)
        traffic[wl] = tot
    ll = os.path.join(src, "launches_%s.csv" % wl)
    if os.path.exists(ll):
        rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
        hdr = None; out_rows = []
        for r in rows:
            if "Kernel Name" in r:
                hdr = r; continue
            if hdr and len(r) == len(hdr):
                dd = dict(zip(hdr, r))
                if dd.get("Metric Name") == "gpu__time_duration.sum":
                    out_rows.append((dd["ID"], dd["Kernel Name"][:80], dd["Metric Value"], dd["Metric Unit"]))
        with open(os.path.join(dst, "%s_%s_launches.csv" % (tag, wl)), "w") as fh:
            fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)\nid,kernel,duration,unit\n")
            for r in out_rows:
                fh.write(",".join('"%s"' % x for x in r) + "\n")
if "c4_step" in traffic and "c4" in traffic:
    traffic["c4_map_kernel"] = traffic["c4"]
    traffic["c4"] = traffic["c4"] + traffic.pop("c4_step")      # one env-step = step kernel + map kernel
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(dst)))
