#!/bin/bash
# 2-GPU call r3b: fused peer optimizer test, N = 2 bench line (learner_dp with the three-stream backward, circle_config4), stage-2 data-parallel training smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_peer_gpu.py -m gpu -q -s > gpurun_out/r3b_pytest_peer.log 2>&1; echo "exit $?" >> gpurun_out/r3b_pytest_peer.log; tail -15 gpurun_out/r3b_pytest_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 200 --warmup 20 --e2e-steps 5 > gpurun_out/r3b_bench_n2.json 2> gpurun_out/r3b_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r3b_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], json.dumps(d['learner_dp'], indent=1)); print(json.dumps(d.get('circle_config4')))"; tail -5 gpurun_out/r3b_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    ppo_stage2.py --num-worlds 6 --updates 2 --policy-path gpurun_out/r3b_policy > gpurun_out/r3b_train_stage2_n2.log 2>&1
echo "stage2 dp exit $?" >> gpurun_out/r3b_train_stage2_n2.log; tail -4 gpurun_out/r3b_train_stage2_n2.log; grep -i "peer" log/*/output.log 2>/dev/null | tail -2
rm -rf gpurun_out/r3b_policy
