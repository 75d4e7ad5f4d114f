#!/bin/bash
# GPU call r2a: full -m gpu suite, a bench line, ncu DRAM traffic of >=130 consecutive ring launches of the tick,
# ncu --set full captures of raycast-alone / GAE / Adam / PPO loss.
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.txt 2>&1; tail -3 gpurun_out/r2a_smoke.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:rlca_world_kernel -s 30 -c 140 --csv --log-file gpurun_out/r2a_tick_traffic.csv \
    python bench.py --steps 200 --warmup 10 --no-cpu --e2e-steps 2 > gpurun_out/r2a_ncu_bench.log 2>&1
tail -3 gpurun_out/r2a_tick_traffic.csv
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'rlca_world_kernel|gae_kernel|adam_kernel|ppo_loss_kernel' -c 14 -o gpurun_out/r2a_targets \
    python tools/profile_targets.py > gpurun_out/r2a_targets.log 2>&1
tail -2 gpurun_out/r2a_targets.log
ls -la gpurun_out
