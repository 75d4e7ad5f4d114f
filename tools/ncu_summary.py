"""Text summary of an .ncu-rep (one block per kernel launch, the metrics the profiles/ summaries quote):
python tools/ncu_summary.py gpurun_out/x.ncu-rep "header line" > profiles/x_ncu_summary.txt"""
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'dram__bytes_read.sum',
           'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
rep = sys.argv[1]
print('# ' + (sys.argv[2] if len(sys.argv) > 2 else rep))
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--metrics', ','.join(METRICS)], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('== ' + r[hdr.index('Kernel Name')])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print('   %-98s %16s %s' % (m, r[i], units[i]))
