# quick A/B of an on-chip min-sum variant: the on-chip tests, then the three min-sum workloads
python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for wl in I80_nmsa_q030 A79_nmsa_q020 I80_nmsa_q015; do
  python bench.py --workload $wl --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2g_$wl.json 2> gpurun_out/r2g_$wl.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2g_$wl.json')); p=d['roofline'].get('phases'); print('$wl value %.4f'%d['value'], 'cn %.2f vn %.2f batch %.2f'%(p['check_ms'],p['variable_ms'],p['batch_ms']))
except Exception as e: print('$wl failed', e); print(open('gpurun_out/r2g_$wl.err').read()[-1500:])
"
done
