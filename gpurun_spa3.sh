for v in 1 2 4; do
python bench.py --workload A82_spa_q0162 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --frames-per-lane $v > gpurun_out/spa3_v$v.json 2>> gpurun_out/spa3.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/spa3_v$v.json')); r=d['roofline']
    print('SPA V=$v: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', round(r.get('sched_ms_per_step'),1), 'tiles', d['config']['pool_tiles'])
except Exception as e: print('V=$v failed', e)
PY
done
tail -3 gpurun_out/spa3.err
