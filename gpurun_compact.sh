timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
for wl in A82_spa_q0162 A82_spalin_q0162; do
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/cp_${wl}.json 2>> gpurun_out/cp.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/cp_${wl}.json')); r=d['roofline']
    print('$wl: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'fer', d['config']['fer'], 'launches', d['gpu_launches'])
except Exception as e: print('$wl failed', e)
PY
done
python bench.py --workload A79_nmsa_q020 --path 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/cp_a79.json 2>> gpurun_out/cp.err; python -c "
import json; d=json.load(open('gpurun_out/cp_a79.json')); print('A79 streaming: %.3f Gbit/s'%d['value'])"
python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/cp_l100k.json 2>> gpurun_out/cp.err; python -c "
import json; d=json.load(open('gpurun_out/cp_l100k.json')); print('L100k streaming: %.3f Gbit/s'%d['value'])"
tail -3 gpurun_out/cp.err
