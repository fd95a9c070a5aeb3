# round 2, call Y: ncu --set full of the float32 on-chip min-sum kernel with 8-byte records (I80 NMSA @ 3 %)
CMD="python bench.py --workload I80_nmsa_q030 --record-bytes 8 --frames 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
$CMD > gpurun_out/r2y_plain.json 2> gpurun_out/r2y_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum" -s 1 -c 1 -o gpurun_out/prof_r02y_rec8 $CMD > gpurun_out/r2y_ncu.log 2>&1
tail -2 gpurun_out/r2y_ncu.log
