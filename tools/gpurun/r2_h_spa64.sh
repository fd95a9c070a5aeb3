# round 2, call H: branch-free double tanh / atanh in the float64 sum-product check node -- parity (full suite incl. the two
# new operating-point tests) and throughput of the float64 streaming SPA path before / after (before: r02_b_*_before.json)
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2h_pytest.txt; cat gpurun_out/r2h_pytest.txt
python -m pytest tests/test_gpu_points.py tests/test_gpu_large.py -m gpu -q -s 2>&1 | grep -E "precision|alg=0|passed|failed" | head -20
python bench.py --workload A82_spa_q0162 --precision 64 --frames 8192 --no-cpu-baseline --no-secondary > gpurun_out/r2h_A82_spa_f64.json 2> gpurun_out/r2h_A82_spa_f64.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_A82_spa_f64.json')); print('A82 SPA f64 value %.3f e2e %.3f it %.2f'%(d['value'], d['e2e']['value'], d['config']['mean_iterations_executed']), d['roofline']['both_kernels'])"
python bench.py --workload L100k_spa_q084 --frames 1024 --no-cpu-baseline --no-secondary > gpurun_out/r2h_L100k_spa.json 2> gpurun_out/r2h_L100k_spa.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_L100k_spa.json')); print('L100k SPA f64 value %.3f e2e %.3f it %.2f'%(d['value'], d['e2e']['value'], d['config']['mean_iterations_executed']), d['roofline']['both_kernels'])"
