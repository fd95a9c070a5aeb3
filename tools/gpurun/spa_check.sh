# SPA streaming path: parity tests, then the SPA / SPA-lin workloads at 1, 2 and 4 frames per lane
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q 2>&1 | tail -4
for wl in A82_spa_q0162 A82_spalin_q0162; do
for fpl in 1 2 4; do
python bench.py --workload $wl --frames-per-lane $fpl --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2> gpurun_out/spa_$fpl.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl fpl=$fpl %.3f Gbit/s frac %.3f'%(d['value'], d['roofline']['frac']))"
done
done
