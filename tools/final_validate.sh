#!/bin/bash
# Round-end validation in one gpurun call: sanitizer over every kernel path, full GPU test suite, smoke, the bench
# line, and the ncu launch list (durations + DRAM bytes) of the same bench command.
mkdir -p gpurun_out
( echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | tail -45
  echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_run.py 2>&1 | tail -8 ) > gpurun_out/sanitizer.log 2>&1
python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 900 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/sanitizer.log
