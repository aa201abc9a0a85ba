timeout 300 python tools/warp_dev.py --modes=warp32,warp --pmodes=warp32,warp 2>&1 | tee gpurun_out/w32_dev3.log | grep -v " ok$" | tail -20
for env in "X=1" "CTC_B200_FULL_GRIDS=1"; do
echo "== $env"
env $env ALT_MODE=warp32 python tools/warp_alt_time.py "50-200,100-140" default mix1 2>&1 | tee -a gpurun_out/w32_alt4.log
env $env ALT_MODE=warp python tools/warp_alt_time.py "50-200" default 2>&1 | tee -a gpurun_out/w32_alt4.log
done
