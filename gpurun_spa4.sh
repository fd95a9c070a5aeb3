timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q -k "spa or A82-0 or single or K1_3 or K1_4 or K1_hi" 2>&1 | tail -8
for v in 1 4; do
python bench.py --workload A82_spa_q0162 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --frames-per-lane $v > gpurun_out/spa4_v$v.json 2>> gpurun_out/spa4.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/spa4_v$v.json')); r=d['roofline']
    print('SPA V=$v: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'fer', d['config']['fer'])
except Exception as e: print('V=$v failed', e)
PY
done
