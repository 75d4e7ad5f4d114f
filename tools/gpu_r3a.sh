#!/bin/bash
# GPU call r3a: stage-2 group barrier scanned by the lanes of the re-spawn warp; all env parity tests; tick split
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py tests/test_eval_gpu.py -m gpu -q > gpurun_out/r3a_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3a_pytest_env.log
tail -3 gpurun_out/r3a_pytest_env.log
timeout 300 python tools/exp_tick_split.py > gpurun_out/r3a_tick_split.jsonl 2>&1; cat gpurun_out/r3a_tick_split.jsonl
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'physics|lidar' --launch-skip 40 --launch-count 2 -o gpurun_out/r3a_stage2 -f python tools/profile_scenario.py stage2 94 2 40 > gpurun_out/r3a_ncu.log 2>&1; tail -1 gpurun_out/r3a_ncu.log
