timeout 1500 python -m pytest tests/test_gpu_onchip.py tests/test_gpu_keygen.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
for wl in I80_nmsa_q030 A79_nmsa_q020 I80_nmsa_q015; do
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl %.3f Gbit/s'%d['value'])"
done
