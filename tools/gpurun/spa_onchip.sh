# on-chip sum-product kernel: parity tests, then SPA / SPA-lin throughput on both paths
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_onchip.py -x -q 2>&1 | tail -6
for wl in A82_spa_q0162 A82_spalin_q0162; do
for path in 1 2; do
python bench.py --workload $wl --path $path --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> gpurun_out/spa_p$path.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl path=$path %.3f Gbit/s frac %.3f threads %s'%(d['value'], d['roofline']['frac'], d['config'].get('onchip_threads')))"
done
done
tail -3 gpurun_out/spa_p2.err
