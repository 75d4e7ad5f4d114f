#!/bin/bash
# N = 8 (and 4) check of the bench line incl. the learner_dp section (fused peer-memory optimizer kernel over 8 GPUs)
mkdir -p gpurun_out
for n in 8 4; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
      bench.py --gpus $n --steps 500 --warmup 50 --e2e-steps 50 > gpurun_out/r2q_bench_n$n.json 2> gpurun_out/r2q_bench_n$n.err
  echo "n=$n exit $?"
  python -c "
import json
d=json.loads(open('gpurun_out/r2q_bench_n$n.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['learner_dp'], indent=1))"
  tail -3 gpurun_out/r2q_bench_n$n.err
done
