for opt in "" "--no-compaction" "--pool-slots 4096" "--pool-slots 4096 --no-compaction"; do
python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $opt > gpurun_out/l100k.json 2>> gpurun_out/l100k.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/l100k.json')); r=d['roofline']
    print('L100k [$opt]: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', round(r.get('sched_ms_per_step'),1), 'tiles', d['config']['pool_tiles'], 'launches', d['gpu_launches'])
except Exception as e: print('[$opt] failed', e)
PY
done
tail -3 gpurun_out/l100k.err
