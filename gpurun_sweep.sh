timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 1 2 4; do
  timeout 300 python bench.py --frames 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --frames-per-lane $v > gpurun_out/sw_v$v.json 2>>gpurun_out/sw.err
  python - <<PY
import json
d=json.load(open('gpurun_out/sw_v$v.json'))
r=d['roofline']
print('V=$v value %.3f Gbit/s  step_frac %.3f cn %.3f vn %.3f sched_ms %.2f'%(d['value'], r['whole_step_frac'], r['both_kernels']['cn']['frac'], r['both_kernels']['vn']['frac'], r['sched_ms_per_step']))
PY
done
for ps in 4096 8192 16384; do
  timeout 300 python bench.py --workload A79_nmsa_q020 --frames 32768 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --pool-slots $ps > gpurun_out/sw_p$ps.json 2>>gpurun_out/sw.err
  python - <<PY
import json
d=json.load(open('gpurun_out/sw_p$ps.json'))
r=d['roofline']
print('A79 pool=$ps value %.3f Gbit/s  step_frac %.3f cn %.3f vn %.3f sched_ms %.2f ms/step %.1f launches %d'%(d['value'], r['whole_step_frac'], r['both_kernels']['cn']['frac'], r['both_kernels']['vn']['frac'], r['sched_ms_per_step'], d['ms_per_step'], d['gpu_launches']))
PY
done
tail -5 gpurun_out/sw.err
