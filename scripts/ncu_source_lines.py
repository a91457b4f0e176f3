"""Stall samples per CUDA source line from an .ncu-rep:  python scripts/ncu_source_lines.py <rep> <kernel regex> [launch index] [top n]
(reads `ncu --page source --print-source cuda,sass --csv`; the rows with a line number carry that line's totals)."""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split per kernel instance on the header rows
starts = [i for i, x in enumerate(rows) if len(x) > 5 and x[0] == "Line No"]
starts.append(len(rows))
blk = rows[starts[which]:starts[which + 1]]
hdr = blk[0]
si = hdr.index("# Samples")
lines = [x for x in blk[1:] if len(x) > si and x[0].strip().isdigit()]
tot = sum(int(x[si] or 0) for x in lines)
print("kernel instances:", len(starts) - 1, " total samples:", tot)
for x in sorted(lines, key=lambda x: -int(x[si] or 0))[:top]:
    print(f"{x[0]:>5s} {int(x[si]):6d} {100 * int(x[si]) / max(tot, 1):5.1f}%  {x[1].strip()[:120]}")
