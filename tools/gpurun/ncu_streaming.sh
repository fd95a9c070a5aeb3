set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -c 600 gpurun_out/bench_r01c.json
CMD="python bench.py --frames 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel|sched_kernel" -s 150 -c 8 -o gpurun_out/prof_r01c $CMD > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -5
