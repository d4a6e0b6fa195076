"""Summarise an .ncu-rep: headline metrics and the hottest SASS lines.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_top]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu', 'sm__inst_executed_pipe_lsu.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct',
        'smsp__warps_active.avg.per_cycle_active', 'launch__registers_per_thread ',
        'launch__occupancy_limit', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'smsp__average_warps_issue_stalled', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed_op_shared', 'sm__inst_executed_pipe_fp64.sum ']
for h, u, v in zip(hdr, units, vals):
    hh = h + ' '
    if any(k in hh for k in keys):
        try:
            if 'stalled' in h and float(v.replace(',', '')) < 0.05:
                continue
        except ValueError:
            pass
        print('%-88s %-10s %s' % (h, u, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(int(r[ix['# Samples']] or 0) for r in data) or 1
tot_i = sum(int(r[ix['Instructions Executed']] or 0) for r in data) or 1
print('total samples %d, warp instructions %d' % (tot_s, tot_i))
top = sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:n_top]
for r in top:
    print('%6.2f%% samp %6.2f%% inst thr=%5s  %s' % (
        100 * int(r[ix['# Samples']]) / tot_s, 100 * int(r[ix['Instructions Executed']]) / tot_i,
        r[ix['Avg. Threads Executed']][:5], r[ix['Source']][:90]))
