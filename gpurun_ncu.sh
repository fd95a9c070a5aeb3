CMD="python bench.py --frames 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel|sched_kernel" -s 150 -c 5 -o gpurun_out/prof_r01b $CMD > gpurun_out/ncu_f.log 2>&1
CMD2="python bench.py --workload A79_nmsa_q020 --frames 16384 --pool-slots 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel|sched_kernel" -s 60 -c 3 -o gpurun_out/prof_r01b_a79 $CMD2 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out | tail -8
