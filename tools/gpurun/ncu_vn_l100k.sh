# ncu --set full of the streaming VN / CN kernels on the n = 102400 code (L100k NMSA @ QBER 6 %)
CMD="python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_l100k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"vn_kernel|cn_kernel" -s 40 -c 4 -o gpurun_out/prof_l100k $CMD > gpurun_out/ncu_l100k.log 2>&1
tail -2 gpurun_out/ncu_l100k.log; tail -c 600 gpurun_out/plain_l100k.log
