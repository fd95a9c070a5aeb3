python tools/config_parity.py --config adaptiveR --ref-trials 50 --full --out gpurun_out/adaptiveR_parity.json > /dev/null 2> gpurun_out/cfgAR.err; grep -v WARNING gpurun_out/cfgAR.err | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/adaptiveR_parity.json')); r=d['runs']
print('ref cpu %.1fs (%d trials)'%(r['reference_cpu']['seconds'], d['ref_trials']))
for k in ('qkdldpc_sim_default','qkdldpc_sim_fp32','qkdldpc_sim_fp64'):
    rows=r[k]['rows']
    print(' ',k,'%.1fs'%r[k]['seconds'],'combos',len(rows),'csv identical',r[k]['csv_identical'],'rows identical',sum(x['row_identical'] for x in rows),'fer inside ci',sum(x['fer_inside_ci'] for x in rows), 'max |iter mean diff| %.2f'%max(abs(x['iter_mean_ref']-x['iter_mean_gpu']) for x in rows))
f=r['qkdldpc_sim_full']; print('  full: %.1fs for %d combinations x %d trials; reference extrapolated %.0fs'%(f['seconds'],f['combinations'],f['trials'],f['reference_cpu_seconds_extrapolated']))
PY
