# round 2, call K: evidence for the float32 on-chip min-sum kernel on its new layout and for the default-policy configs
# (1) ncu --set full at the bench's frame count (32 768 frames x 100 iterations in ONE launch), (2) launch list of the default
# bench command, (3) ADAPTIVE R.json through qkdldpc_sim in the default precision policy vs the reference executable.
CMD="python bench.py --frames 32768 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
$CMD > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum" -s 1 -c 1 -o gpurun_out/prof_r02k_onchip_32768 $CMD > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log; ls -la gpurun_out/prof_r02k_onchip_32768.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_default_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_ncu_launch.log 2>&1
tail -2 gpurun_out/r2k_ncu_launch.log | cut -c1-200
python tools/config_parity.py --config adaptiveR --ref-trials 50 --full --out gpurun_out/r2k_adaptiveR_parity.json > /dev/null 2> gpurun_out/r2k_cfgAR.err; grep -v WARNING gpurun_out/r2k_cfgAR.err | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_adaptiveR_parity.json')); r=d['runs']
print('ref cpu %.1fs (%d trials)'%(r['reference_cpu']['seconds'], d['ref_trials']))
for k in ('qkdldpc_sim_default','qkdldpc_sim_fp32','qkdldpc_sim_fp64'):
    rows=r[k]['rows']
    print(' ',k,'%.1fs'%r[k]['seconds'],'combos',len(rows),'csv identical',r[k]['csv_identical'],'rows identical',sum(x['row_identical'] for x in rows),'fer inside ci',sum(x['fer_inside_ci'] for x in rows), 'max |iter mean diff| %.2f'%max(abs(x['iter_mean_ref']-x['iter_mean_gpu']) for x in rows))
f=r['qkdldpc_sim_full']; print('  full: %.1fs for %d combinations x %d trials; reference extrapolated %.0fs'%(f['seconds'],f['combinations'],f['trials'],f['reference_cpu_seconds_extrapolated']))
PY
