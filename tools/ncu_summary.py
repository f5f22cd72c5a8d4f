"""Condense an `ncu --set full` report into the per-kernel summary CSV kept under profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>_summary.csv"""
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'launch__grid_size', 'launch__registers_per_thread',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
cols = [h.index('Kernel Name')] + [h.index(m) for m in METRICS if m in h]
w = csv.writer(sys.stdout)
w.writerow([h[c] for c in cols])
w.writerow([rows[1][c] for c in cols])
for r in rows[2:]:
    w.writerow([r[c][:70] for c in cols])
