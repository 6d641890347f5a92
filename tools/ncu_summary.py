"""Key metrics + stall breakdown of an ncu report: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "")[:70], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor"]
    for k in keys:
        if k in d:
            print("  %-62s %s %s" % (k, d[k], u.get(k, "")))
    st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v)
          for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k}
    print("  stalls (warps per issue):", ", ".join("%s %.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
