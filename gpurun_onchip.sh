timeout 900 python -m pytest tests/test_gpu_onchip.py -x -q 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_onchip.py 2>&1 | tail -8
for path in 2 1; do
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --path $path > gpurun_out/oc_bench_p$path.json 2> gpurun_out/oc_bench_p$path.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/oc_bench_p$path.json')); r=d['roofline']
    print('path $path: value %.3f Gbit/s e2e %.3f ms/step %.1f kernel %s frac %.3f whole %.3f launches %d clocks %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel'], r['frac'], r['whole_step_frac'], d['gpu_launches'], d['clocks']))
except Exception as e: print('path $path failed', e); print(open('gpurun_out/oc_bench_p$path.err').read()[-2000:])
PY
done
for wl in A79_nmsa_q020 I80_nmsa_q015; do for path in 2 1; do
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --path $path > gpurun_out/oc_${wl}_p$path.json 2>> gpurun_out/oc.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/oc_${wl}_p$path.json')); r=d['roofline']
    print('$wl path $path: value %.3f Gbit/s ms/step %.1f frac %.3f mean it %.2f fer %.4f'%(d['value'], d['ms_per_step'], r['whole_step_frac'], d['config']['mean_iterations_executed'], d['config']['fer']))
except Exception as e: print('$wl path $path failed', e)
PY
done; done
for t in 256 384 512; do
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --path 2 --onchip-threads $t --frames 16384 > gpurun_out/oc_t$t.json 2>> gpurun_out/oc.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/oc_t$t.json')); print('threads $t: value %.3f Gbit/s'%d['value'], d['config']['pool_tiles'])
except Exception as e: print('threads $t failed', e)
PY
done
tail -5 gpurun_out/oc.err
