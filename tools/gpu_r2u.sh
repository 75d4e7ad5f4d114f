#!/bin/bash
# GPU call r2u: learner backward on three streams + 32-deep fp32 GEMM tiles + one-launch minibatch gather;
# big-map lidar with the start cell's distance in shared memory, 6 CTAs/SM, one viewer per CTA by default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learner_gpu.py tests/test_eval_gpu.py -m gpu -q -x > gpurun_out/r2u_pytest_learner.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2u_pytest_learner.log
tail -5 gpurun_out/r2u_pytest_learner.log
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q -k "circle or big or global" > gpurun_out/r2u_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2u_pytest_env.log
tail -3 gpurun_out/r2u_pytest_env.log
RLCA_BWD_STREAMS=0 timeout 300 python tools/exp_learner_step.py > gpurun_out/r2u_learner.jsonl 2>&1
timeout 300 python tools/exp_learner_step.py >> gpurun_out/r2u_learner.jsonl 2>&1; cat gpurun_out/r2u_learner.jsonl
timeout 300 python tools/exp_tick_split.py circle > gpurun_out/r2u_tick_split.jsonl 2>&1; cat gpurun_out/r2u_tick_split.jsonl
timeout 300 python tools/exp_circle_shape.py > gpurun_out/r2u_circle_shape.jsonl 2>&1; cat gpurun_out/r2u_circle_shape.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_launches_learner.csv python tools/exp_learner_step.py > /dev/null 2>&1
