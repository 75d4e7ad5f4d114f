#!/bin/bash
# 2-GPU call: the bench line at N=2 (learner_dp section: overlapped gradient all-reduce) and a stage-2 data-parallel
# training smoke (per-rank filter_index -> agreed minibatch schedule, must not hang)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 300 --warmup 20 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.err
tail -c 1500 gpurun_out/r2l_bench_n2.json; tail -5 gpurun_out/r2l_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    ppo_stage2.py --num-worlds 6 --updates 2 --policy-path gpurun_out/r2l_policy > gpurun_out/r2l_train_stage2_n2.log 2>&1
echo "stage2 dp exit $?" >> gpurun_out/r2l_train_stage2_n2.log; tail -6 gpurun_out/r2l_train_stage2_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    ppo_stage1.py --num-worlds 8 --updates 2 --policy-path gpurun_out/r2l_policy1 > gpurun_out/r2l_train_stage1_n2.log 2>&1
echo "stage1 dp exit $?" >> gpurun_out/r2l_train_stage1_n2.log; tail -4 gpurun_out/r2l_train_stage1_n2.log
rm -rf gpurun_out/r2l_policy gpurun_out/r2l_policy1
