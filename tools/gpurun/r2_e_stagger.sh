# round 2, call E: do the CTAs that share an SM run their check (ALU-bound) and variable (LSU-bound) phases in lockstep? Phase
# clocks of the kernel, and the headline with the CTAs of the 2nd / 3rd resident wave started late by a fraction of an iteration.
for ns in 0 15000 31000 46000; do
  QKDLDPC_OC_STAGGER_NS=$ns python bench.py --workload I80_nmsa_q030 --frames 16384 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2e_s$ns.json 2> gpurun_out/r2e_s$ns.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2e_s$ns.json')); print('stagger $ns value %.4f'%d['value'], d['roofline'].get('phases'))
except Exception as e: print('stagger $ns failed', e); print(open('gpurun_out/r2e_s$ns.err').read()[-1500:])
"
done
for wl in I80_nmsa_q015 A79_nmsa_q020; do
for ns in 0 31000; do
  QKDLDPC_OC_STAGGER_NS=$ns python bench.py --workload $wl --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2e_${wl}_s$ns.json 2> gpurun_out/r2e_${wl}_s$ns.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2e_${wl}_s$ns.json')); print('$wl stagger $ns value %.4f'%d['value'], d['roofline'].get('phases'))
except Exception as e: print('$wl stagger $ns failed', e)
"
done
done
