"""Summarise an ncu report: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__warps_eligible.avg.per_cycle_active']
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print('%-75s %-8s %s' % (k, units[i], [r[i] for r in data]))
print('--- warp stall reasons (per issue active) > 0.3')
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        vals = [r[i] for r in data]
        try:
            if max(float(v) for v in vals) > 0.3:
                print('%-40s %s' % (h.replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', ''), vals))
        except ValueError:
            pass
