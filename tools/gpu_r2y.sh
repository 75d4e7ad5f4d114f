#!/bin/bash
# GPU call r2y: same-box A/B of the distance-field jump (first cell at chessboard distance d  vs  d steps), both with the
# float-reciprocal floor division; parity of both
mkdir -p gpurun_out
L=rl_collision_avoidance_b200
for v in chess steps chess steps; do
  if [ $v = steps ]; then cp $L/librlca.so /tmp/librlca_chess.so; cp $L/librlca_steps.so $L/librlca.so; fi
  if [ $v = chess ] && [ -f /tmp/librlca_chess.so ]; then cp /tmp/librlca_chess.so $L/librlca.so; fi
  echo "== $v" >> gpurun_out/r2y_ab.jsonl
  timeout 300 python tools/exp_tick_split.py circle >> gpurun_out/r2y_ab.jsonl 2>&1
  if [ ! -f gpurun_out/r2y_pytest_$v.log ]; then
    timeout 900 python -m pytest tests/test_env_parity_gpu.py tests/test_env_fullsize_gpu.py -m gpu -q -k "circle or big or global" > gpurun_out/r2y_pytest_$v.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2y_pytest_$v.log
    tail -2 gpurun_out/r2y_pytest_$v.log
  fi
done
cat gpurun_out/r2y_ab.jsonl
cp /tmp/librlca_chess.so $L/librlca.so
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'big_lidar' --launch-skip 30 --launch-count 1 -o gpurun_out/r2y_circle -f python tools/profile_scenario.py circle 41 1 40 > gpurun_out/r2y_ncu.log 2>&1; tail -1 gpurun_out/r2y_ncu.log
