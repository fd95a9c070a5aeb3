# round 2, call R: where the float64 on-chip min-sum kernel spends its time -- phase clocks, then ncu --set full on one launch
for spec in "I80_nmsa_q030 64" "A82_aomsa_q0161 0"; do
  set -- $spec
  python bench.py --workload $1 --precision $2 --frames 32768 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2r_$1_p$2.json 2> gpurun_out/r2r_$1_p$2.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2r_$1_p$2.json')); p=d['roofline'].get('phases') or {}; print('$1 precision $2: value %.4f'%d['value'], d['dtype'], 'cn %.2f vn %.2f batch %.2f'%(p.get('check_ms',0),p.get('variable_ms',0),p.get('batch_ms',0)))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/r2r_$1_p$2.err').read()[-1500:])
"
done
CMD="python bench.py --workload I80_nmsa_q030 --precision 64 --frames 4736 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
ncu --set full --clock-control none --import-source on -k regex:"onchip_minsum64" -s 1 -c 1 -o gpurun_out/prof_r02r_onchip64 $CMD > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log
