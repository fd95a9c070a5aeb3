# round 2, call S: software-pipelined index loads in both on-chip min-sum kernels -- tests, then A/B against the float32
# kernel without them (libqkdldpc_cuda_f32base.so.variant = the same tree with the previous onchip_minsum.cuh)
python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2s_pytest.txt
run() {  # tag workload precision
  python bench.py --workload $2 --precision $3 --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2s_$1_$2_p$3.json 2> gpurun_out/r2s_$1_$2_p$3.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2s_$1_$2_p$3.json')); p=d['roofline'].get('phases') or {}; print('$1 $2 precision $3: value %.4f'%d['value'], d['dtype'], 'cn %.2f vn %.2f batch %.2f'%(p.get('check_ms',0),p.get('variable_ms',0),p.get('batch_ms',0)))
except Exception as e: print('$1 $2 failed', e); print(open('gpurun_out/r2s_$1_$2_p$3.err').read()[-1500:])
"
}
run pipe I80_nmsa_q030 0; run pipe A79_nmsa_q020 0; run pipe I80_nmsa_q015 0; run pipe I80_nmsa_q030 64; run pipe A82_aomsa_q0161 0; run pipe I80_nmsa_q030 0
cp qkd_ldpc_v_b200/libqkdldpc_cuda.so /tmp/main.so; cp qkd_ldpc_v_b200/libqkdldpc_cuda_f32base.so.variant qkd_ldpc_v_b200/libqkdldpc_cuda.so
run base I80_nmsa_q030 0; run base A79_nmsa_q020 0; run base I80_nmsa_q015 0; run base I80_nmsa_q030 0
cp /tmp/main.so qkd_ldpc_v_b200/libqkdldpc_cuda.so
