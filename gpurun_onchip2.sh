timeout 900 python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -5
for wl in I80_nmsa_q030 A79_nmsa_q020 I80_nmsa_q015; do
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --path 2 > gpurun_out/oc2_${wl}.json 2>> gpurun_out/oc2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/oc2_${wl}.json')); r=d['roofline']
    print('$wl: value %.3f Gbit/s ms/step %.1f frac %.3f mean it %.2f fer %.4f ctas %d clocks %s'%(d['value'], d['ms_per_step'], r['whole_step_frac'], d['config']['mean_iterations_executed'], d['config']['fer'], d['config']['pool_tiles'], d['clocks']))
except Exception as e: print('$wl failed', e)
PY
done
CMD="python bench.py --frames 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --path 2"
$CMD > gpurun_out/plain_oc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum" -s 1 -c 1 -o gpurun_out/prof_r01e_onchip $CMD > gpurun_out/ncu_oc.log 2>&1
tail -2 gpurun_out/ncu_oc.log; tail -3 gpurun_out/oc2.err
