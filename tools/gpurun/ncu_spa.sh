# on-chip sum-product kernel: CTA-size sweep, then one ncu --set full capture (A82 SPA @ QBER 1.62 %)
for t in 512 768 896 960 1024; do
python bench.py --workload A82_spa_q0162 --path 2 --onchip-threads $t --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> gpurun_out/spa_t.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('threads=$t %.3f Gbit/s'%d['value'])"
done
CMD="python bench.py --workload A82_spa_q0162 --frames 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --path 2"
$CMD > gpurun_out/plain_spa.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_spa" -s 1 -c 1 -o gpurun_out/prof_r01j_onchip_spa $CMD > gpurun_out/ncu_spa.log 2>&1
tail -3 gpurun_out/ncu_spa.log; ls -la gpurun_out/prof_r01j_onchip_spa.ncu-rep
