timeout 1500 python -m pytest tests/test_gpu_keygen.py tests/test_sim_csv.py tests/test_gpu_onchip.py -x -q 2>&1 | tail -4
python tools/config_parity.py --config adaptiveR --ref-trials 50 --full --out gpurun_out/adaptiveR_parity2.json > /dev/null 2> gpurun_out/cfgAR2.err; grep -v WARNING gpurun_out/cfgAR2.err | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/adaptiveR_parity2.json')); r=d['runs']
print('ref cpu %.1fs (%d trials)'%(r['reference_cpu']['seconds'], d['ref_trials']))
for k in ('qkdldpc_sim_fp32','qkdldpc_sim_fp64'):
    rows=r[k]['rows']
    print(' ',k,'%.1fs'%r[k]['seconds'],'combos',len(rows),'csv identical',r[k]['csv_identical'],'rows identical',sum(x['row_identical'] for x in rows),'fer inside ci',sum(x['fer_inside_ci'] for x in rows))
f=r['qkdldpc_sim_full']; print('  full: %.1fs for %d combinations x %d trials; reference extrapolated %.0fs'%(f['seconds'],f['combinations'],f['trials'],f['reference_cpu_seconds_extrapolated']))
PY
