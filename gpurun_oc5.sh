timeout 900 python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -3
for wl in I80_nmsa_q030 A79_nmsa_q020 I80_nmsa_q015; do
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/oc5_${wl}.json 2>> gpurun_out/oc5.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/oc5_${wl}.json')); r=d['roofline']
    print('$wl: value %.3f Gbit/s ms/step %.1f frac %.3f ctas %d'%(d['value'], d['ms_per_step'], r['whole_step_frac'], d['config']['pool_tiles']))
except Exception as e: print('$wl failed', e)
PY
done
tail -3 gpurun_out/oc5.err
