"""Summarise an .ncu-rep: key metrics per kernel + top stalled SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for d in data:
    print('=' * 100); print(d[idx['Kernel Name']][:110])
    for w in want:
        if w in idx: print(f"  {w:82s} {d[idx[w]][:18]:>18s} {units[idx[w]]}")
    s = sorted(((float(d[idx[h]] or 0), h) for h in stalls), reverse=True)[:6]
    print('  stalls/issue:', [(round(v, 2), h.split('stalled_')[1].replace('_per_issue_active.ratio', '')) for v, h in s])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
for sec in src.split('"Kernel Name",')[1:]:
    lines = sec.split('\n'); r = list(csv.reader(lines[1:])); h = r[0]; rs = [x for x in r[1:] if len(x) == len(h)]
    si, so = h.index('# Samples'), h.index('Source'); tot = sum(int(x[si] or 0) for x in rs) or 1
    print('-' * 100); print(lines[0][:100], 'samples', tot)
    for x in sorted(rs, key=lambda x: -int(x[si] or 0))[:topn]:
        st = {k: int(x[h.index(k)] or 0) for k in h if k.startswith('stall_') and 'Not Issued' not in k}
        m = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"   {100*int(x[si])/tot:5.1f}%  {x[so][:78]:78s} {m}")
