# round 2, call Q: float64 on-chip min-sum kernel on the tables of onchip_layout.hpp (16-byte record + c2 array, sign tests on
# the FP64 pipe, clamp after the minimum) -- FP64 / ALU pipe rates first, then the on-chip and parity tests, then A/B numbers
./tools/microbench/fp64_pipe > gpurun_out/r2q_fp64_pipe.txt 2>&1; cat gpurun_out/r2q_fp64_pipe.txt
python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2q_pytest.txt
for spec in "A82_aomsa_q0161 0" "I80_nmsa_q030 64" "A79_nmsa_q020 64" "I80_nmsa_q030 0"; do
  set -- $spec
  python bench.py --workload $1 --precision $2 --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2q_$1_p$2.json 2> gpurun_out/r2q_$1_p$2.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2q_$1_p$2.json')); p=d['roofline'].get('phases') or {}; print('$1 precision $2: value %.4f'%d['value'], d['dtype'], d['config'].get('decoder_path'), 'cn %.2f vn %.2f batch %.2f'%(p.get('check_ms',0),p.get('variable_ms',0),p.get('batch_ms',0)))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/r2q_$1_p$2.err').read()[-1500:])
"
done
