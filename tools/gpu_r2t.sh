#!/bin/bash
# GPU call r2t: big-map lidar v3 (per-edge scatter items, long inverse lists flattened over the warp)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q -k "circle or big or global" > gpurun_out/r2t_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest_env.log
tail -4 gpurun_out/r2t_pytest_env.log
timeout 300 python tools/exp_tick_split.py circle > gpurun_out/r2t_tick_split.jsonl 2>&1; cat gpurun_out/r2t_tick_split.jsonl
timeout 300 python tools/exp_circle_shape.py > gpurun_out/r2t_circle_shape.jsonl 2>&1; cat gpurun_out/r2t_circle_shape.jsonl
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'big_lidar' --launch-skip 30 --launch-count 1 -o gpurun_out/r2t_circle -f python tools/profile_scenario.py circle 41 1 40 > gpurun_out/r2t_ncu.log 2>&1; tail -2 gpurun_out/r2t_ncu.log
