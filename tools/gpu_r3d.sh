#!/bin/bash
# 1-GPU validation call r3d (final kernels of the round): whole -m gpu suite, smoke, bench line, reference arm, sanitizer (env + learner), ncu of the tick and of circle.world, traffic, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3d_pytest.log; tail -5 gpurun_out/r3d_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3d_smoke.txt 2>&1; tail -3 gpurun_out/r3d_smoke.txt
timeout 600 python bench.py > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; tail -c 300 gpurun_out/r3d_bench.json; tail -3 gpurun_out/r3d_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r3d_bench_reference.json 2>> gpurun_out/r3d_bench.err; cat gpurun_out/r3d_bench_reference.json | cut -c1-400
OUT=gpurun_out/r3d_sanitizer.txt
echo "# compute-sanitizer on tools/sanitize_smoke.py --env-only --circle (final round-2 kernels)" > $OUT
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> $OUT
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_smoke.py --env-only --circle 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize workload|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" | head -30 >> $OUT
done
echo "## memcheck, learner (three-stream backward, one-launch gather)" >> $OUT
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize workload|ERROR SUMMARY|Error|error" | head -30 >> $OUT
cat $OUT
timeout 300 python tools/exp_tick_split.py > gpurun_out/r3d_tick_split.jsonl 2>&1; cat gpurun_out/r3d_tick_split.jsonl
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:'big_lidar|physics' --launch-skip 60 --launch-count 2 -o gpurun_out/r3d_circle -f python tools/profile_scenario.py circle 41 1 40 > gpurun_out/r3d_ncu_circle.log 2>&1; tail -1 gpurun_out/r3d_ncu_circle.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rlca_physics_kernel|rlca_lidar_kernel' -s 80 -c 4 -o gpurun_out/r3d_tick \
    python bench.py --steps 100 --warmup 10 --no-cpu --no-sections --no-graph --e2e-steps 2 > gpurun_out/r3d_ncu.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none \
    -k regex:'rlca_physics_kernel|rlca_lidar_kernel' -s 60 -c 280 --csv --log-file gpurun_out/r3d_tick_traffic.csv \
    python bench.py --steps 200 --warmup 10 --no-cpu --no-sections --no-graph --e2e-steps 2 > gpurun_out/r3d_ncu_traffic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r3d_launches_bench.csv \
    python bench.py --steps 100 --warmup 5 --no-cpu --no-sections --no-graph --e2e-steps 3 > gpurun_out/r3d_launches.log 2>&1
ls -la gpurun_out | grep r3d
