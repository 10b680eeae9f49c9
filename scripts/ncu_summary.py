"""Print the metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg.per_second', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for vals in rows[2:]:
    print("----")
    d = dict(zip(hdr, vals))
    for h in hdr:
        if h in want:
            print(f"{h:70s} {d[h]:>16s} {units[hdr.index(h)]}")
    stalls = [(h, float(d[h].replace(',', ''))) for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued') and d[h]]
    tot = sum(v for _, v in stalls) or 1
    for h, v in sorted(stalls, key=lambda t: -t[1])[:10]:
        print(f"   stall {h.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {100*v/tot:5.1f}%")
