#!/bin/bash
# GPU call r3k: training-loop tests (trainer, checkpoints, resume) on the optimizer step of r3j
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_training_gpu.py -m gpu -q -x > gpurun_out/r3k_pytest_training.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3k_pytest_training.log
tail -5 gpurun_out/r3k_pytest_training.log
