#!/bin/bash
# GPU call r2z: big-map lidar with robots behind the viewer dropped at the pair stage, 128-bit fill of hit[]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q -k "circle or big or global" > gpurun_out/r2z_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2z_pytest_env.log
tail -3 gpurun_out/r2z_pytest_env.log
timeout 300 python tools/exp_tick_split.py circle > gpurun_out/r2z_tick_split.jsonl 2>&1; cat gpurun_out/r2z_tick_split.jsonl
timeout 300 python tools/exp_tick_split.py circle >> gpurun_out/r2z_tick_split.jsonl 2>&1; tail -1 gpurun_out/r2z_tick_split.jsonl
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'big_lidar' --launch-skip 30 --launch-count 1 -o gpurun_out/r2z_circle -f python tools/profile_scenario.py circle 41 1 40 > gpurun_out/r2z_ncu.log 2>&1; tail -1 gpurun_out/r2z_ncu.log
