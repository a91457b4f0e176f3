"""gpurun_out/traffic.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum) ->
profiles/r02_traffic.json: per kernel family, DRAM bytes per launch averaged over the launches of one forward."""
import csv, json, collections, sys
src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/traffic.csv"
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
fam = {"dwconv_kernel": "dwconv", "stem_tc_kernel": "stem"}
acc = collections.defaultdict(lambda: collections.defaultdict(float))
for r in rows:
    name, metric, unit, val = r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    key = None
    for k, v in fam.items():
        if k in name: key = v
    if "pw_gemm_tc_kernel<1" in name or "pw_gemm_tc_kernel<(bool)1" in name: key = "project_gemm"
    elif "pw_gemm_tc_kernel" in name: key = "expand_gemm"
    if key is None: continue
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
    acc[key][metric] += val * mult
    acc[key]["n:" + metric] += 1
out = {}
for k, d in acc.items():
    n = d["n:dram__bytes_read.sum"] or 1
    out[k] = {"launches": int(n), "dram_bytes_per_launch": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / n,
              "dram_read_bytes_total": d["dram__bytes_read.sum"], "dram_write_bytes_total": d["dram__bytes_write.sum"],
              "ncu_time_us_total": d["gpu__time_duration.sum"], "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one batch-256 forward"}
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
