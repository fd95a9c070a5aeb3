timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_keygen.py tests/test_gpu_large.py -x -q 2>&1 | tail -3
for wl in I80_nmsa_q030 A79_nmsa_q020 A82_spa_q0162; do
python bench.py --workload $wl --path 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/vn_$wl.json 2>> gpurun_out/vn.err
python - <<PY
import json
d=json.load(open('gpurun_out/vn_$wl.json')); r=d['roofline']
print('$wl streaming: value %.3f Gbit/s ms/step %.1f'%(d['value'], d['ms_per_step']), {k:(round(v['ms_per_step'],1), round(v['frac'],3)) for k,v in r.get('both_kernels',{}).items()}, 'sched', round(r.get('sched_ms_per_step'),1))
PY
done
python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/vn_l100k.json 2>> gpurun_out/vn.err; python -c "
import json; d=json.load(open('gpurun_out/vn_l100k.json')); print('L100k streaming: %.3f Gbit/s'%d['value'], d['roofline']['both_kernels'])"
tail -3 gpurun_out/vn.err
