"""Top SASS instructions of an ncu report by stall samples, with their neighbours.
usage: python tools/ncu_sass.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; ins = []
for r in rows:
    if not r: continue
    if "Source" in r and "# Samples" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            ins.append((int(d.get("# Samples") or 0), int(d.get("Instructions Executed") or 0), d.get("Source", "").strip(), d))
        except ValueError:
            pass
tot = sum(i[0] for i in ins) or 1
print("sass instructions %d, samples %d" % (len(ins), tot))
order = sorted(range(len(ins)), key=lambda k: -ins[k][0])[:top]
stall_cols = [c for c in (hdr or []) if c.startswith("stall_")]
for k in order:
    s, n, src, d = ins[k]
    st = sorted(((float(d.get(c) or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% exec %9d  [%d] %s   %s" % (s, 100.0 * s / tot, n, k, src[:110], " ".join("%s=%g" % (c, v) for v, c in st if v)))
