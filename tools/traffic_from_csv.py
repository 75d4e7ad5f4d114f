"""DRAM traffic per launch of the tick kernel from an ncu metrics CSV of consecutive ring launches
(ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:rlca_world_kernel -s 30 -c 140
 python bench.py --steps 200 --no-graph ...)  ->  profiles/r2_step_kernel_traffic.json (read by bench.py)."""
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if 'Metric Name' in r][0]
h = rows[hi]
kn, mn, mu, mv, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value'), h.index('ID')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1, 'us': 1e3, 'ms': 1e6}
per = {}
for r in rows[hi + 1:]:
    if len(r) <= mv or 'rlca_world_kernel<0' not in r[kn].replace('(int)', '').replace(' ', ''):
        continue
    per.setdefault(r[idc], {})[r[mn]] = float(r[mv].replace(',', '')) * scale.get(r[mu], 1)
n = len(per)
rd = sum(v.get('dram__bytes_read.sum', 0) for v in per.values()) / n
wr = sum(v.get('dram__bytes_write.sum', 0) for v in per.values()) / n
ns = sum(v.get('gpu__time_duration.sum', 0) for v in per.values()) / n
out = {'dram_bytes_per_launch': rd + wr, 'dram_read_bytes_per_launch': rd, 'dram_write_bytes_per_launch': wr,
       'launches_averaged': n, 'avg_ns_under_ncu': ns, 'source': src,
       'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over consecutive launches of the headline tick '
              '(171 worlds x 24 robots x 512 beams) writing a 128-slot obs ring (1.08 GB > L2)'}
json.dump(out, open(dst, 'w'), indent=1)
print(out)
