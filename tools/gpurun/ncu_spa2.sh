# one ncu --set full capture of the on-chip sum-product kernel (A82 SPA @ QBER 1.62 %, 4736 frames)
CMD="python bench.py --workload A82_spa_q0162 --frames 4736 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --path 2 $SPA_EXTRA"
$CMD > gpurun_out/plain_spa.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_spa" -s 1 -c 1 -o gpurun_out/prof_spa_iter $CMD > gpurun_out/ncu_spa.log 2>&1
tail -2 gpurun_out/ncu_spa.log; ls -la gpurun_out/prof_spa_iter.ncu-rep
