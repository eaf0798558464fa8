"""Dev helper: print key metrics of every kernel in an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in idx:
            print(f'{w:72s} {r[idx[w]]} {units[idx[w]]}')
