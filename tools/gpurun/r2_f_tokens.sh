# round 2, call F: (1) phase times with 1, 2, 3 CTAs per SM; (2) at most T CTAs of an SM inside the variable phase at a time
for fr in 148 296 444 888; do
  python bench.py --workload I80_nmsa_q030 --frames $fr --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2f_f$fr.json 2> gpurun_out/r2f_f$fr.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2f_f$fr.json')); p=d['roofline'].get('phases'); print('frames $fr value %.4f ms %.3f'%(d['value'], d['ms_per_step']), 'cn %.3f vn %.3f batch %.3f'%(p['check_ms'],p['variable_ms'],p['batch_ms']))
except Exception as e: print('frames $fr failed', e); print(open('gpurun_out/r2f_f$fr.err').read()[-1500:])
"
done
for tk in 1 2; do
for wl in I80_nmsa_q030 I80_nmsa_q015 A79_nmsa_q020; do
  QKDLDPC_OC_VN_TOKENS=$tk python bench.py --workload $wl --frames 16384 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2f_${wl}_t$tk.json 2> gpurun_out/r2f_${wl}_t$tk.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2f_${wl}_t$tk.json')); p=d['roofline'].get('phases'); print('$wl tokens $tk value %.4f'%d['value'], 'cn %.3f vn %.3f batch %.3f'%(p['check_ms'],p['variable_ms'],p['batch_ms']))
except Exception as e: print('$wl tokens $tk failed', e); print(open('gpurun_out/r2f_${wl}_t$tk.err').read()[-1500:])
"
done
done
