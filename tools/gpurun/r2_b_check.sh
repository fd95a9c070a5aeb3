# round 2, call B: full GPU test-suite (final keys, ADAPTIVE T.json CSV, pipelined host batches), the default bench line with
# its `workloads` array, the reference arm, and the float64 sum-product throughput (streaming) before any rewrite.
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.txt
cat gpurun_out/r2b_pytest.txt
( time python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2b_bench_default.json'))
    print('HEAD value %.4f e2e %.4f frac %.2f launches %d cpu %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline'] and d['cpu_baseline']['value']))
    for w in d['workloads']:
        print('%-18s value %.3f e2e %.3f dtype %s it %.2f fer %.4f frac %.2f wsf %.2f %s' % (w['workload_id'], w['value'], w['e2e']['value'], w['dtype'], w['mean_iterations_executed'], w['fer'], w['roofline']['frac'], w['roofline']['whole_step_frac'], w['decoder_path'][:12]))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2b_bench_default.err').read()[-3000:])
PY
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err; cat gpurun_out/r2b_bench_reference.json | cut -c1-300
python bench.py --workload A82_spa_q0162 --precision 64 --frames 8192 --no-cpu-baseline --no-secondary > gpurun_out/r2b_A82_spa_f64.json 2> gpurun_out/r2b_A82_spa_f64.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_A82_spa_f64.json')); print('A82 SPA f64 value %.3f e2e %.3f it %.2f'%(d['value'], d['e2e']['value'], d['config']['mean_iterations_executed']), d['roofline']['both_kernels'])"
python bench.py --workload A82_spa_q0162 --path 1 --frames 32768 --no-cpu-baseline --no-secondary > gpurun_out/r2b_A82_spa_f32_streaming.json 2> gpurun_out/r2b_A82_spa_f32s.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_A82_spa_f32_streaming.json')); print('A82 SPA f32 streaming value %.3f e2e %.3f'%(d['value'], d['e2e']['value']), d['roofline']['both_kernels'])"
