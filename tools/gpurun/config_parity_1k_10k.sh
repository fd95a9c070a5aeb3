python tools/config_parity.py --config config1k --ref-trials 20000 --full --out gpurun_out/config1k_parity.json > /dev/null 2> gpurun_out/cfg1k.err; tail -2 gpurun_out/cfg1k.err
python tools/config_parity.py --config config10k --ref-trials 5000 --full --out gpurun_out/config10k_parity.json > /dev/null 2> gpurun_out/cfg10k.err; tail -2 gpurun_out/cfg10k.err
python - <<'PY'
import json
for c in ('config1k','config10k'):
    try:
        d=json.load(open(f'gpurun_out/{c}_parity.json'))
    except Exception as e:
        print(c,'failed',e); continue
    r=d['runs']
    print(c,'ref cpu %.1fs (%d trials)'%(r['reference_cpu']['seconds'], d['ref_trials']))
    for k in ('qkdldpc_sim_fp32','qkdldpc_sim_fp64'):
        print(' ',k,'%.1fs'%r[k]['seconds'],'csv identical',r[k]['csv_identical'], [(x['fer_ref'],x['fer_gpu'],x['fer_inside_ci'],x['iter_mean_ref'],x['iter_mean_gpu']) for x in r[k]['rows']])
    f=r.get('qkdldpc_sim_full')
    if f: print('  full: %.1fs for %d trials/matrix; reference extrapolated %.0fs; rows'%(f['seconds'],f['trials'],f['reference_cpu_seconds_extrapolated']), [(x['fer'],x['iter_mean']) for x in f['rows']], [round(x['decoded_gbit_s'],2) for x in f['sidecar']])
PY
