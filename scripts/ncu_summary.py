"""Summarise an .ncu-rep (raw page) for the stage kernel: time, DRAM traffic, pipe utilisation, stalls."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index('Kernel Name')]}
    for w in want:
        if w in hdr:
            d[w] = r[hdr.index(w)] + " " + units[hdr.index(w)]
    st = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
            try:
                st.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    d["stalls_pct"] = {h: round(v / tot * 100, 1) for v, h in sorted(st, reverse=True)[:8]}
    res.append(d)
print(json.dumps(res, indent=1))
