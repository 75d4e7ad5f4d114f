#!/bin/bash
# GPU call r3f: small-map lidar with 128-bit hit[] fill and list copy (list cells 16-byte aligned behind a 4-word header)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py tests/test_eval_gpu.py -m gpu -q > gpurun_out/r3f_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3f_pytest_env.log
tail -3 gpurun_out/r3f_pytest_env.log
timeout 300 python tools/exp_tick_split.py stage1 stage2 > gpurun_out/r3f_tick_split.jsonl 2>&1
timeout 300 python tools/exp_tick_split.py stage1 stage2 >> gpurun_out/r3f_tick_split.jsonl 2>&1; cat gpurun_out/r3f_tick_split.jsonl
timeout 300 python bench.py --steps 1000 --warmup 50 --no-cpu --no-sections --e2e-steps 50 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step']*1e3, d['value']/1e6, d['e2e']['value']/1e6)"
