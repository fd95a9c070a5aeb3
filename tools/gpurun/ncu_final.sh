CMD="python bench.py --frames 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --path 2"
$CMD > gpurun_out/plain_oc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum" -s 1 -c 1 -o gpurun_out/prof_r01i_onchip $CMD > gpurun_out/ncu_oc.log 2>&1
CMD2="python bench.py --workload A82_spa_q0162 --frames 8192 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/plain_spa.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cn_kernel|vn_kernel|sched_kernel|sched_move" -s 40 -c 8 -o gpurun_out/prof_r01i_spa $CMD2 > gpurun_out/ncu_spa.log 2>&1
CMD3="python tools/keygen_bench.py"
$CMD3 > gpurun_out/plain_kg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ref_keygen" -s 1 -c 1 -o gpurun_out/prof_r01i_keygen $CMD3 > gpurun_out/ncu_kg.log 2>&1
tail -2 gpurun_out/ncu_oc.log gpurun_out/ncu_spa.log gpurun_out/ncu_kg.log; cat gpurun_out/plain_kg.log | tail -3
