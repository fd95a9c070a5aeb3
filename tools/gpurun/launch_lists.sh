# ncu launch lists (gpu__time_duration.sum per launch) of the default bench command and of the SPA workload
for wl in I80_nmsa_q030 A82_spa_q0162; do
CMD="python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$wl.json 2> gpurun_out/plain_$wl.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$wl.csv $CMD > gpurun_out/ncu_ll_$wl.log 2>&1
tail -c 300 gpurun_out/plain_$wl.json; echo; grep -c onchip gpurun_out/launches_$wl.csv
done
