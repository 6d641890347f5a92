"""Top source lines of an ncu report by stall samples / executed instructions.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections
def num(v):
    try:
        return int(v)
    except (TypeError, ValueError):
        return 0


rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
SORT = 1 if (len(sys.argv) > 3 and sys.argv[3] == "inst") else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None; hdr = None; agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += num(d.get("# Samples"))
        a[1] += num(d.get("Instructions Executed"))
        a[2] += 1
tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, total warp-instr %d (over all captured launches)" % (tot_s, tot_i))
print("%-16s %5s %7s %7s  %s" % ("file", "line", "samp%", "inst%", "source"))
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
    print("%-16s %5d %6.1f%% %6.1f%%  %s" % (f, ln, 100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, src))
