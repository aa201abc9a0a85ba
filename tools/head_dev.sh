#!/bin/bash
# One gpurun call while developing the classifier-head kernels: parity tests, timings, optionally (argument "ncu") one
# ncu pass over the head kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 2>&1 | tail -15 | tee gpurun_out/head_pytest.log
timeout 300 python tools/head_bench.py 2>&1 | tee gpurun_out/head_bench.log
for a in "$@"; do
  if [ "$a" = ncu ]; then
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:head_ -s 9 -c 9 -o gpurun_out/head_ncu -f python tools/head_profile_one.py 256 29 > gpurun_out/head_ncu.log 2>&1
    tail -3 gpurun_out/head_ncu.log
  fi
done
