# round 2, call X: 8-byte records in the float32 on-chip min-sum kernel -- tests, then A/B against the 16-byte records
python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_random_codes.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2x_pytest.txt
run() {  # tag workload record-bytes
  python bench.py --workload $2 --record-bytes $3 --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2x_$1_$2.json 2> gpurun_out/r2x_$1_$2.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2x_$1_$2.json')); p=d['roofline'].get('phases') or {}; print('$1 $2: value %.4f'%d['value'], d['dtype'], 'record bytes', d['config']['onchip_record_bytes'], 'threads', d['config']['onchip_threads'], 'cn %.2f vn %.2f batch %.2f'%(p.get('check_ms',0),p.get('variable_ms',0),p.get('batch_ms',0)))
except Exception as e: print('$1 $2 failed', e); print(open('gpurun_out/r2x_$1_$2.err').read()[-1500:])
"
}
for wl in I80_nmsa_q030 A79_nmsa_q020 I80_nmsa_q015; do run rec8 $wl 8; run rec16 $wl 16; done
