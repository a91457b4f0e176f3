"""Per-kernel stall-sample breakdown of an .ncu-rep by code region: prints every SASS line with >= N samples in address order
(argv: report, launch index, min samples) -- shows which warp role (producer / MMA / transform / epilogue) waits where."""
import csv, subprocess, sys, io
rep, idx, mins = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name'] + [len(allrows)]
rows = allrows[starts[idx]:starts[idx + 1]]
print(rows[0][1][:120])
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
si, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[si] or 0) for r in data); print('samples', tot)
for i, r in enumerate(data):
    n = int(r[si] or 0)
    if n >= mins or any(k in r[1] for k in ('UTMALDG', 'UTCHMMA', 'UTMASTG', 'LDTM')):
        st = {k: int(r[hdr.index(k)] or 0) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
        m = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {n:5d} {100*n/tot:5.1f}% {r[ie]:>8s}  {r[1].strip()[:70]:70s} {m}")
