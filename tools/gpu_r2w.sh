#!/bin/bash
# GPU call r2w: conv tower backward with the next sample's operands prefetched; weight-image prep of the backward on a
# side stream; Wc prep only for the CUDA-core conv path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_learner_gpu.py tests/test_eval_gpu.py -m gpu -q -x > gpurun_out/r2w_pytest_learner.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest_learner.log
tail -5 gpurun_out/r2w_pytest_learner.log
timeout 300 python tools/exp_learner_step.py > gpurun_out/r2w_learner.jsonl 2>&1; cat gpurun_out/r2w_learner.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_launches_learner.csv python tools/exp_learner_step.py > /dev/null 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
