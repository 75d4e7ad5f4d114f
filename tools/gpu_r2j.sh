#!/bin/bash
# GPU call r2j: env kernels v3 (table-driven lidar, window collision test, big-map split): parity first, then numbers.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q > gpurun_out/r2j_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest_env.log
tail -15 gpurun_out/r2j_pytest_env.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_env_parity_gpu.py --deselect tests/test_env_fullsize_gpu.py > gpurun_out/r2j_pytest_rest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest_rest.log
tail -8 gpurun_out/r2j_pytest_rest.log
timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 600 gpurun_out/r2j_bench.json; tail -5 gpurun_out/r2j_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rlca_physics_kernel|rlca_lidar_kernel' -s 80 -c 4 -o gpurun_out/r2j_tick \
    python bench.py --steps 100 --warmup 10 --no-cpu --no-sections --no-graph --e2e-steps 2 > gpurun_out/r2j_ncu.log 2>&1
tail -2 gpurun_out/r2j_ncu.log
ls -la gpurun_out | grep r2j
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
    -k regex:'rlca_physics_kernel|rlca_lidar_kernel' -s 60 -c 280 --csv --log-file gpurun_out/r2j_tick_traffic.csv \
    python bench.py --steps 200 --warmup 10 --no-cpu --no-sections --no-graph --e2e-steps 2 > gpurun_out/r2j_ncu_traffic.log 2>&1
tail -2 gpurun_out/r2j_tick_traffic.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rlca_physics_kernel|rlca_big_lidar_kernel' -s 12 -c 2 -o gpurun_out/r2j_circle \
    python tools/profile_scenario.py circle 41 1 12 > gpurun_out/r2j_ncu_circle.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rlca_physics_kernel|rlca_lidar_kernel' -s 12 -c 2 -o gpurun_out/r2j_stage2 \
    python tools/profile_scenario.py stage2 94 2 12 > gpurun_out/r2j_ncu_stage2.log 2>&1
ls -la gpurun_out | grep r2j
