#!/bin/bash
# GPU call r3j: optimizer step that also writes the tf32 split of the fc1 weights (rlca_policy_adam_step)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_learner_gpu.py tests/test_eval_gpu.py -m gpu -q -x > gpurun_out/r3j_pytest_learner.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3j_pytest_learner.log
tail -5 gpurun_out/r3j_pytest_learner.log
timeout 300 python tools/exp_learner_step.py > gpurun_out/r3j_learner.jsonl 2>&1; cat gpurun_out/r3j_learner.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
