# round 2, call V: frame set-up / read-out of the on-chip min-sum kernels with their L2 / DRAM latencies overlapped (four index
# blocks, four slot words, four read-out rounds in flight; Alice's words fetched up front) -- tests, then A/B against the
# previous build (libqkdldpc_cuda_base.so.variant)
python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_random_codes.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2v_pytest.txt
run() {  # tag workload precision
  python bench.py --workload $2 --precision $3 --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2v_$1_$2_p$3.json 2> gpurun_out/r2v_$1_$2_p$3.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2v_$1_$2_p$3.json')); p=d['roofline'].get('phases') or {}; print('$1 $2 precision $3: value %.4f'%d['value'], d['dtype'], 'cn %.2f vn %.2f batch %.2f'%(p.get('check_ms',0),p.get('variable_ms',0),p.get('batch_ms',0)))
except Exception as e: print('$1 $2 failed', e); print(open('gpurun_out/r2v_$1_$2_p$3.err').read()[-1500:])
"
}
run new A79_nmsa_q020 0; run new I80_nmsa_q015 0; run new I80_nmsa_q030 0; run new A82_aomsa_q0161 0; run new A79_nmsa_q020 0
cp qkd_ldpc_v_b200/libqkdldpc_cuda.so /tmp/main.so; cp qkd_ldpc_v_b200/libqkdldpc_cuda_base.so.variant qkd_ldpc_v_b200/libqkdldpc_cuda.so
run base A79_nmsa_q020 0; run base I80_nmsa_q015 0; run base I80_nmsa_q030 0; run base A82_aomsa_q0161 0; run base A79_nmsa_q020 0
cp /tmp/main.so qkd_ldpc_v_b200/libqkdldpc_cuda.so
