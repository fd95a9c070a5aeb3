timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_keygen.py -x -q 2>&1 | tail -3
python bench.py --workload A79_nmsa_q020 --path 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('A79 streaming %.3f'%d['value'], d['roofline']['both_kernels']['vn']['frac'])"
