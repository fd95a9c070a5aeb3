# quick check of the on-chip sum-product kernel: parity subset, then throughput
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "spa or fp32_messages" 2>&1 | tail -3
for wl in A82_spa_q0162 A82_spalin_q0162; do
python bench.py --workload $wl --path 2 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e $SPA_EXTRA 2> gpurun_out/spa_q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl %.3f Gbit/s threads %s'%(d['value'], d['config'].get('onchip_threads')))"
done
