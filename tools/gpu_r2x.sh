#!/bin/bash
# GPU call r2x: distance-field walks jump to the first cell at chessboard distance d (not d steps); hand-placed wall /
# corner / outside poses test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q -k "circle or big or global" > gpurun_out/r2x_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2x_pytest_env.log
tail -8 gpurun_out/r2x_pytest_env.log
timeout 300 python tools/exp_tick_split.py circle > gpurun_out/r2x_tick_split.jsonl 2>&1; cat gpurun_out/r2x_tick_split.jsonl
timeout 300 python tools/exp_circle_shape.py > gpurun_out/r2x_circle_shape.jsonl 2>&1; cat gpurun_out/r2x_circle_shape.jsonl
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'big_lidar' --launch-skip 30 --launch-count 1 -o gpurun_out/r2x_circle -f python tools/profile_scenario.py circle 41 1 40 > gpurun_out/r2x_ncu.log 2>&1; tail -1 gpurun_out/r2x_ncu.log
