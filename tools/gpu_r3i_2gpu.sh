#!/bin/bash
# 2-GPU call r3i: N = 2 bench line with the learner_dp variants measured in an order that does not leak the overlap mode
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 2 --steps 200 --warmup 20 --e2e-steps 5 > gpurun_out/r3i_bench_n2.json 2> gpurun_out/r3i_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r3i_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], json.dumps(d['learner_dp'], indent=1)); print(json.dumps(d.get('circle_config4')))"; tail -3 gpurun_out/r3i_bench_n2.err
