timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q -k "spa or SPA or A82-0 or A82-1 or single or operating" 2>&1 | tail -12
for wl in A82_spa_q0162 A82_spalin_q0162 A79_nmsa_q020; do
extra=""; if [ $wl = A79_nmsa_q020 ]; then extra="--path 1"; fi
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $extra > gpurun_out/spa2_${wl}.json 2>> gpurun_out/spa2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/spa2_${wl}.json')); r=d['roofline']
    print('$wl: value %.3f Gbit/s ms/step %.1f whole %.3f mean it %.2f fer %.4f launches %d'%(d['value'], d['ms_per_step'], r['whole_step_frac'], d['config']['mean_iterations_executed'], d['config']['fer'], d['gpu_launches']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', r.get('sched_ms_per_step'), 'tiles', d['config']['pool_tiles'])
except Exception as e: print('$wl failed', e)
PY
done
tail -3 gpurun_out/spa2.err
