#!/bin/bash
# 4-GPU call r3g: N = 4 bench line (weak-scaling tick, learner_dp after the three-stream backward, circle_config4)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 4 --steps 200 --warmup 20 --e2e-steps 5 > gpurun_out/r3g_bench_n4.json 2> gpurun_out/r3g_bench_n4.err
python -c "
import json
d=json.loads(open('gpurun_out/r3g_bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], json.dumps(d['learner_dp'], indent=1)); print(json.dumps(d.get('circle_config4')))"; tail -3 gpurun_out/r3g_bench_n4.err
