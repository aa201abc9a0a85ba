python tools/w32_case.py 5079 warp32 2>&1 | grep "b=13"
timeout 300 python tools/warp_dev.py --pmodes=warp32 --modes=warp32 2>&1 | grep -v " ok$" | tee gpurun_out/w32_dev4.log | head -14
timeout 900 python tools/fuzz.py 400 7000 2>&1 | tail -6 | tee gpurun_out/fuzz_w32c.log
python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 900 -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_w32b.log
