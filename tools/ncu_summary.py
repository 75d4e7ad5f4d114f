"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + hottest SASS regions."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_bytes.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic', 'smsp__average_warp_latency_issue_stalled', 'sm__cycles_elapsed.max']
for i, h in enumerate(hdr):
    if h in want:
        print(f'{h:70s}', [r[i] for r in rows[2:5]], rows[1][i])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
sec = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) > 6:
        sec.append(r)
h = rows[1]
ie, ss = h.index('Instructions Executed'), h.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[ie] or 0) for r in sec)
tots = sum(int(r[ss] or 0) for r in sec)
print('total warp-instructions', tot, 'stall samples', tots)
for i, r in enumerate(sec):
    c = int(r[ie] or 0)
    if c > tot * top or int(r[ss] or 0) > tots * top * 2:
        print(f'{i:5d} {r[1][:64]:64s} inst {c:9d} {100*c/tot:5.1f}%  stall {100*int(r[ss] or 0)/max(tots,1):5.1f}%')
