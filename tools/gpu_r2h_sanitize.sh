#!/bin/bash
# compute-sanitizer on the round-2 env kernels (physics, table lidar, big-map lidar): memcheck, racecheck, synccheck
mkdir -p gpurun_out
OUT=gpurun_out/r2h_sanitizer.txt
echo "# compute-sanitizer on tools/sanitize_smoke.py --env-only --circle (ticks stage1 / stage2 / 180 beams with the scan FIFO, raycast, rlca_env_step_host in its three traffic modes, circle.world; round-2 kernels: rlca_physics_kernel, rlca_lidar_kernel, rlca_big_lidar_kernel)" > $OUT
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> $OUT
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_smoke.py --env-only --circle 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize workload|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -30 >> $OUT
done
cat $OUT
timeout 300 python tools/exp_tick_split.py > gpurun_out/r2h_tick_split.jsonl 2>&1; cat gpurun_out/r2h_tick_split.jsonl
