# L2-residency experiment: does a message pool that fits in the 126 MB L2 beat the HBM-streaming pool?
for v in 4 1; do
for ps in 128 256 384 512 1024 4096; do
  timeout 300 python bench.py --frames 8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --frames-per-lane $v --pool-slots $ps > gpurun_out/l2_v${v}_p$ps.json 2>>gpurun_out/l2.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/l2_v${v}_p$ps.json'))
    r=d['roofline']
    print('V=$v pool=$ps value %.3f Gbit/s  step_frac %.3f cn %.3f vn %.3f sched_ms %.2f ms/step %.1f pool_bytes %.1f MB'%(d['value'], r['whole_step_frac'], r['both_kernels']['cn']['frac'], r['both_kernels']['vn']['frac'], r['sched_ms_per_step'], d['ms_per_step'], d['config']['pool_bytes']/1e6))
except Exception as e: print('V=$v pool=$ps failed', e)
PY
done
done
tail -5 gpurun_out/l2.err
