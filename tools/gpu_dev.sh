#!/bin/bash
# One gpurun call during development: parity tests, pipe micro-benchmarks, timing sweep.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
./build/ubench 2>&1 | tee gpurun_out/ubench.log
timeout 600 python tools/sweep.py "$@" 2>&1 | tee gpurun_out/sweep.log
