# round 2, call D: ncu --set full of the float32 on-chip min-sum kernel on the new layout (I80 NMSA @3 %, 1184 frames x 100 iterations)
CMD="python bench.py --frames 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --path 2"
$CMD > gpurun_out/r2d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum" -s 1 -c 1 -o gpurun_out/prof_r02d_onchip $CMD > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/prof_r02d_onchip.ncu-rep
