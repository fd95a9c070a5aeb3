# round 2: full verification (GPU tests, smoke, default bench, reference arm, config 100k.json CSV parity); files r2ba_*
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2ba_pytest.txt; cat gpurun_out/r2ba_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r2ba_bench_default.json 2> gpurun_out/r2ba_bench_default.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2ba_bench_default.json'))
    print('HEAD value %.4f e2e %.4f frac %.2f launches %d cpu %s clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks']))
    for w in d['workloads']:
        print('%-18s value %.3f e2e %.3f dtype %s it %.2f fer %.4f frac %.2f wsf %.2f %s %s' % (w['workload_id'], w['value'], w['e2e']['value'], w['dtype'], w['mean_iterations_executed'], w['fer'], w['roofline']['frac'], w['roofline']['whole_step_frac'], w['decoder_path'][:12], w.get('streaming')))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/r2ba_bench_default.err').read()[-3000:])
PY
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2ba_bench_reference.json 2> gpurun_out/r2ba_bench_reference.err; cut -c1-200 gpurun_out/r2ba_bench_reference.json
# config 100k.json as shipped (SPA, default policy float64) through qkdldpc_sim against the reference executable, 100 trials
timeout 400 python tools/config_parity.py --config config100k --ref-trials 100 --out gpurun_out/r2ba_config100k.json > /dev/null 2> gpurun_out/r2ba_config100k.err; tail -2 gpurun_out/r2ba_config100k.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2ba_config100k.json")); r=d["runs"]
    for k,v in r.items(): print(k, {kk: vv for kk, vv in v.items() if kk in ("seconds","trials","csv_identical","threads")})
except Exception as e: print("config100k failed", e)
PY
