# round 2, call M: L2 prefetch of a later item's messages in the narrow variable-node kernel -- distance sweep on the n = 102400 code
for d in 0 2048 8192 16384 32768 65536; do
  QKDLDPC_VN_PREFETCH=$d python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 2 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2m_l100k_pf$d.json 2> gpurun_out/r2m_l100k_pf$d.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2m_l100k_pf$d.json')); b=d['roofline']['both_kernels']; print('L100k prefetch $d value %.3f whole %.3f cn %.3f vn %.3f'%(d['value'], d['roofline']['whole_step_frac'], b['cn']['frac'], b['vn']['frac']))
except Exception as e: print('prefetch $d failed', e); print(open('gpurun_out/r2m_l100k_pf$d.err').read()[-1200:])
"
done
for d in 0 8192; do
  QKDLDPC_VN_PREFETCH=$d python bench.py --workload A79_nmsa_q020 --path 1 --frames 32768 --steps 2 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2m_a79_pf$d.json 2> gpurun_out/r2m_a79_pf$d.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2m_a79_pf$d.json')); b=d['roofline']['both_kernels']; print('A79 streaming prefetch $d value %.3f whole %.3f cn %.3f vn %.3f'%(d['value'], d['roofline']['whole_step_frac'], b['cn']['frac'], b['vn']['frac']))
except Exception as e: print('prefetch $d failed', e)
"
done
