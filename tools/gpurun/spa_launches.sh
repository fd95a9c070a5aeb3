# per-launch durations of the SPA streaming step at 2 and 4 frames per lane (ncu launch list; run after spa_check.sh)
for fpl in 2 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/spa_launches_fpl$fpl.csv \
  python bench.py --workload A82_spa_q0162 --frames-per-lane $fpl --frames 16384 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/spa_ncu_$fpl.log 2>&1
done
