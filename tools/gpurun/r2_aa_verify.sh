# round 2, call AA: full verification after the float64 kernel rebuild and the 8-byte records
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2aa_pytest.txt; cat gpurun_out/r2aa_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r2aa_bench_default.json 2> gpurun_out/r2aa_bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2aa_bench_default.json'))
    print('HEAD value %.4f e2e %.4f frac %.2f launches %d cpu %s clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks']))
    for w in d['workloads']:
        print('%-18s value %.3f e2e %.3f dtype %s it %.2f fer %.4f frac %.2f wsf %.2f %s' % (w['workload_id'], w['value'], w['e2e']['value'], w['dtype'], w['mean_iterations_executed'], w['fer'], w['roofline']['frac'], w['roofline']['whole_step_frac'], w['decoder_path'][:12]))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2aa_bench_default.err').read()[-3000:])
PY
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2aa_bench_reference.json 2> gpurun_out/r2aa_bench_reference.err; cut -c1-200 gpurun_out/r2aa_bench_reference.json
