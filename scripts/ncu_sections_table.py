"""Per-launch table from an ncu sections report: duration, DRAM %, tensor %, occupancy, issue rate, top stalls."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time_us"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs")]
cols = [(m, n) for m, n in cols if m in idx]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
w = csv.writer(sys.stdout)
w.writerow(["#", "kernel"] + [n for _, n in cols] + ["top_stalls"])
for i, d in enumerate(data):
    name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("dfv::", "")
    vals = []
    for m, n in cols:
        v = d[idx[m]].replace(",", "")
        try:
            f = float(v)
            if n == "time_us" and "ms" in rows[1][idx[m]]: f *= 1000.0
            vals.append(f"{f:.1f}" if n.endswith("pct") or n == "time_us" else f"{int(f)}")
        except ValueError:
            vals.append(v)
    s = sorted(((float(d[idx[h]] or 0), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:3]
    w.writerow([i, name[:70]] + vals + [" ".join(f"{k}:{v:.1f}" for v, k in s)])
