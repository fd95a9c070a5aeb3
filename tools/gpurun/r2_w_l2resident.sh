# round 2, call W: does the n = 102400 code stream faster from L2 (126 MB) when the pool is small enough to stay resident?
run() {  # tag, extra flags
  python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 3 --no-cpu-baseline --no-secondary --no-e2e $2 > gpurun_out/r2w_$1.json 2> gpurun_out/r2w_$1.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2w_$1.json')); r=d['roofline']; print('$1: value %.4f'%d['value'], 'pool tiles', d['config'].get('pool_tiles'), 'whole-step frac %.3f'%r['whole_step_frac'], {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items() if k in ('cn_frac','vn_frac','frac')})
except Exception as e: print('$1 failed', e); print(open('gpurun_out/r2w_$1.err').read()[-800:])
"
}
run default ""
run v4_p128 "--pool-slots 128"
run v4_p256 "--pool-slots 256"
run v2_p64 "--pool-slots 64 --frames-per-lane 2"
run v2_p128 "--pool-slots 128 --frames-per-lane 2"
run v1_p64 "--pool-slots 64 --frames-per-lane 1"
