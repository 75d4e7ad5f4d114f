#!/bin/bash
# GPU call r3l: bench line of the final library (no CPU baseline, no extra sections): tick, e2e, learner
mkdir -p gpurun_out
timeout 100 python bench.py --no-cpu --no-sections > gpurun_out/r3l_bench_short.json 2> gpurun_out/r3l_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3l_bench_short.json').read().strip().splitlines()[-1]); print(d['ms_per_step']*1e3, d['value']/1e6, d['e2e']['value']/1e6, d['learner'])"
