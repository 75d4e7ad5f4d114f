#!/bin/bash
# GPU call r2c: env kernels v3 (table-driven lidar, window collision test, big-map split): parity first, then numbers.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q > gpurun_out/r2c_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_env.log
tail -15 gpurun_out/r2c_pytest_env.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_env_parity_gpu.py --deselect tests/test_env_fullsize_gpu.py > gpurun_out/r2c_pytest_rest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_rest.log
tail -8 gpurun_out/r2c_pytest_rest.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 600 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
for s in 1 2 3 4 6 8; do
  timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu --no-sections --e2e-steps 20 --ctas-per-world $s >> gpurun_out/r2c_shapes.jsonl 2>> gpurun_out/r2c_bench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c_shapes.jsonl'):
    d=json.loads(l); print(d['config']['ctas_per_world'], round(d['ms_per_step']*1e3,2),'us', round(d['value']/1e6,1),'M/s')
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rlca_world_kernel -s 40 -c 2 -o gpurun_out/r2c_tick \
    python bench.py --steps 100 --warmup 10 --no-cpu --no-sections --no-graph --e2e-steps 2 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
ls -la gpurun_out | grep r2c
