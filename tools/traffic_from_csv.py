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
    name = r[kn].split('(')[0].replace('void ', '')
    if len(r) <= mv or not name.startswith(('rlca_physics_kernel', 'rlca_lidar_kernel', 'rlca_world_kernel')):
        continue
    per.setdefault((name, r[idc]), {})[r[mn]] = float(r[mv].replace(',', '')) * scale.get(r[mu], 1)
kernels = {}
for (name, _), v in per.items():
    k = kernels.setdefault(name, {'n': 0, 'rd': 0.0, 'wr': 0.0, 'ns': 0.0})
    k['n'] += 1; k['rd'] += v.get('dram__bytes_read.sum', 0); k['wr'] += v.get('dram__bytes_write.sum', 0)
    k['ns'] += v.get('gpu__time_duration.sum', 0)
per_kernel = {name: {'launches_averaged': k['n'], 'dram_read_bytes': k['rd'] / k['n'], 'dram_write_bytes': k['wr'] / k['n'],
                     'avg_ns_under_ncu': k['ns'] / k['n']} for name, k in kernels.items()}
# one tick = one launch of every kernel listed (physics + lidar); the dominant kernel is the lidar
rd = sum(k['dram_read_bytes'] for k in per_kernel.values())
wr = sum(k['dram_write_bytes'] for k in per_kernel.values())
ns = sum(k['avg_ns_under_ncu'] for k in per_kernel.values())
n = min(k['launches_averaged'] for k in per_kernel.values())
out = {'dram_bytes_per_launch': rd + wr, 'dram_read_bytes_per_launch': rd, 'dram_write_bytes_per_launch': wr,
       'launches_averaged': n, 'avg_ns_under_ncu': ns, 'per_kernel': per_kernel, 'source': src,
       'how': 'ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum over consecutive ticks (per tick = physics + lidar launch) of the headline workload '
              '(171 worlds x 24 robots x 512 beams) writing a 128-slot obs ring (1.08 GB > L2)'}
json.dump(out, open(dst, 'w'), indent=1)
print(out)
