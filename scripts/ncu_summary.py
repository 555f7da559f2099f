"""ncu report (--set full) -> small CSV of the metrics the design notes quote.  Usage: python scripts/ncu_summary.py <rep> <out.csv>"""
import csv
import subprocess
import sys

KEYS = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_sample_count']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = [hdr.index(k) for k in KEYS if k in hdr]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
print('wrote', sys.argv[2], len(rows) - 2, 'launches')
