set -x
CMD="python bench.py --frames 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel" -s 120 -c 4 -o gpurun_out/prof_v4 $CMD > gpurun_out/ncu_v4.log 2>&1
$CMD --frames-per-lane 1 > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel" -s 120 -c 4 -o gpurun_out/prof_v1 $CMD --frames-per-lane 1 > gpurun_out/ncu_v1.log 2>&1
tail -3 gpurun_out/ncu_v4.log gpurun_out/ncu_v1.log
ls -la gpurun_out
