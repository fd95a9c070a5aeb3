# round 2, call O: CTA sizes that divide the check phase's groups evenly (I80: 67 groups -> 17 warps x 4 rounds instead of 16 x 5)
for t in 512 544 480; do
  python bench.py --workload I80_nmsa_q030 --onchip-threads $t --frames 32768 --steps 2 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2o_i80_t$t.json 2> gpurun_out/r2o_i80_t$t.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2o_i80_t$t.json')); p=d['roofline']['phases']; print('I80 q030 threads $t value %.4f cn %.1f vn %.1f tiles %s'%(d['value'], p['check_ms'], p['variable_ms'], d['config']['pool_tiles']))
except Exception as e: print('threads $t failed', e); print(open('gpurun_out/r2o_i80_t$t.err').read()[-800:])
"
done
for t in 768 736 704 512; do
  python bench.py --workload A79_nmsa_q020 --onchip-threads $t --frames 65536 --steps 2 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2o_a79_t$t.json 2> gpurun_out/r2o_a79_t$t.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2o_a79_t$t.json')); p=d['roofline']['phases']; print('A79 q020 threads $t value %.4f cn %.1f vn %.1f tiles %s'%(d['value'], p['check_ms'], p['variable_ms'], d['config']['pool_tiles']))
except Exception as e: print('threads $t failed', e)
"
done
for t in 512 544; do
  python bench.py --workload I80_nmsa_q015 --onchip-threads $t --frames 65536 --steps 2 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2o_i80q015_t$t.json 2> gpurun_out/r2o_i80q015_t$t.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2o_i80q015_t$t.json')); print('I80 q015 threads $t value %.4f'%(d['value']))
except Exception as e: print('threads $t failed', e)
"
done
