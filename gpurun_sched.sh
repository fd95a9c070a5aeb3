timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_onchip.py -x -q 2>&1 | tail -5
for opt in "--pool-slots 1664" ""; do
python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $opt > gpurun_out/l100k.json 2>> gpurun_out/l100k.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/l100k.json')); r=d['roofline']
    print('L100k [$opt]: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', round(r.get('sched_ms_per_step'),1), 'tiles', d['config']['pool_tiles'])
except Exception as e: print('[$opt] failed', e)
PY
done
for wl in A82_spa_q0162 I80_nmsa_q030; do
python bench.py --workload $wl --path 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sc_$wl.json 2>> gpurun_out/l100k.err
python - <<PY
import json
d=json.load(open('gpurun_out/sc_$wl.json')); r=d['roofline']
print('$wl streaming: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', round(r.get('sched_ms_per_step'),1))
PY
done
tail -3 gpurun_out/l100k.err
