CMD="python bench.py --frames 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --path 2"
$CMD > gpurun_out/plain_oc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum" -s 1 -c 1 -o gpurun_out/prof_r01d_onchip $CMD > gpurun_out/ncu_oc.log 2>&1
tail -3 gpurun_out/ncu_oc.log; ls -la gpurun_out/prof_r01d_onchip.ncu-rep
