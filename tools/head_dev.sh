#!/bin/bash
# One gpurun call while developing the classifier-head kernels: parity tests, then timings (tensor-core backward, and the
# round-1 FMA backward for comparison).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_head.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout 300 2>&1 | tail -40 | tee gpurun_out/head_pytest.log
timeout 300 python tools/head_bench.py 2>&1 | tee gpurun_out/head_bench.log
CTC_B200_HEAD_FMA=1 timeout 300 python tools/head_bench.py 2>&1 | tee gpurun_out/head_bench_fma.log
