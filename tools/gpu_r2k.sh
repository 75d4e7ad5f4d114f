#!/bin/bash
# GPU call r2k: flat-list scatter restored + uncapped distance field + programmatic dependent launch of the lidar kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q > gpurun_out/r2k_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest_env.log
tail -6 gpurun_out/r2k_pytest_env.log
for pdl in 1 0; do
  RLCA_PDL=$pdl timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu --no-sections --e2e-steps 50 >> gpurun_out/r2k_pdl.jsonl 2>> gpurun_out/r2k_bench.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2k_pdl.jsonl'):
    d=json.loads(l); print(round(d['ms_per_step']*1e3,2),'us', round(d['value']/1e6,1),'M/s', round(d['e2e']['value']/1e6,2))
PY
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_env_parity_gpu.py --deselect tests/test_env_fullsize_gpu.py > gpurun_out/r2k_pytest_rest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2k_pytest_rest.log
tail -4 gpurun_out/r2k_pytest_rest.log
timeout 600 python bench.py > gpurun_out/r2k_bench.json 2>> gpurun_out/r2k_bench.err; tail -c 400 gpurun_out/r2k_bench.json; tail -5 gpurun_out/r2k_bench.err
timeout 300 python tools/exp_tick_split.py > gpurun_out/r2k_tick_split.jsonl 2>&1; cat gpurun_out/r2k_tick_split.jsonl
