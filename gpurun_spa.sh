for wl in A82_spa_q0162 A82_spalin_q0162 L100k_nmsa_q060; do
fr=32768; if [ $wl = L100k_nmsa_q060 ]; then fr=4096; fi
python bench.py --workload $wl --frames $fr --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/spa_${wl}.json 2>> gpurun_out/spa.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/spa_${wl}.json')); r=d['roofline']
    print('$wl: value %.3f Gbit/s ms/step %.1f whole %.3f mean it %.2f fer %.4f launches %d'%(d['value'], d['ms_per_step'], r['whole_step_frac'], d['config']['mean_iterations_executed'], d['config']['fer'], d['gpu_launches']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', r.get('sched_ms_per_step'), 'tiles', d['config']['pool_tiles'])
except Exception as e: print('$wl failed', e)
PY
done
tail -3 gpurun_out/spa.err
