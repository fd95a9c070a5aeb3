# round 2, call T: float64 on-chip min-sum with straight-line variable nodes (dv <= 8) and the 16-bit variable-phase table
python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2t_pytest.txt
run() {  # tag workload precision
  python bench.py --workload $2 --precision $3 --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2t_$1_$2_p$3.json 2> gpurun_out/r2t_$1_$2_p$3.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2t_$1_$2_p$3.json')); p=d['roofline'].get('phases') or {}; print('$1 $2 precision $3: value %.4f'%d['value'], d['dtype'], 'cn %.2f vn %.2f batch %.2f'%(p.get('check_ms',0),p.get('variable_ms',0),p.get('batch_ms',0)))
except Exception as e: print('$1 $2 failed', e); print(open('gpurun_out/r2t_$1_$2_p$3.err').read()[-1500:])
"
}
run a I80_nmsa_q030 64; run a A82_aomsa_q0161 0; run a A79_nmsa_q020 64; run a I80_nmsa_q030 0
