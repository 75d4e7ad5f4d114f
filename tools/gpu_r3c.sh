#!/bin/bash
# GPU call r3c: physics launch with template + distance field of a small map staged in shared memory (cp.async)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py tests/test_eval_gpu.py -m gpu -q > gpurun_out/r3c_pytest_env.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3c_pytest_env.log
tail -3 gpurun_out/r3c_pytest_env.log
timeout 300 python tools/exp_tick_split.py stage1 stage2 > gpurun_out/r3c_tick_split.jsonl 2>&1
timeout 300 python tools/exp_tick_split.py stage1 stage2 >> gpurun_out/r3c_tick_split.jsonl 2>&1; cat gpurun_out/r3c_tick_split.jsonl
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py --env-only 2>&1 | grep -E "sanitize workload|ERROR SUMMARY|Error|error" | head -5
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py --env-only 2>&1 | grep -E "sanitize workload|RACECHECK SUMMARY|hazard" | head -5
