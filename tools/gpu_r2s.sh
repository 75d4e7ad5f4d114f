#!/bin/bash
# GPU call r2s: big-map lidar v2 (compacted in-range pairs, static items first, cooperative list drains)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py tests/test_eval_gpu.py -m gpu -q > gpurun_out/r2s_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_pytest_env.log
tail -6 gpurun_out/r2s_pytest_env.log
timeout 300 python tools/exp_tick_split.py > gpurun_out/r2s_tick_split.jsonl 2>&1; cat gpurun_out/r2s_tick_split.jsonl
timeout 300 python tools/exp_circle_shape.py > gpurun_out/r2s_circle_shape.jsonl 2>&1; cat gpurun_out/r2s_circle_shape.jsonl
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'big_lidar|physics' --launch-skip 60 --launch-count 2 -o gpurun_out/r2s_circle -f python tools/profile_scenario.py circle 41 1 40 > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log
