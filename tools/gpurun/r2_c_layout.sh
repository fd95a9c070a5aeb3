# round 2, call C: the float32 on-chip min-sum kernel on the layout of onchip_layout.hpp (storage order = processing order,
# conflict-aware lanes and edge order): full GPU test-suite, then the headline and the two converging min-sum workloads.
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2c_pytest.txt
cat gpurun_out/r2c_pytest.txt
for wl in I80_nmsa_q030 A79_nmsa_q020 I80_nmsa_q015; do
  python bench.py --workload $wl --no-cpu-baseline --no-secondary > gpurun_out/r2c_$wl.json 2> gpurun_out/r2c_$wl.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2c_$wl.json')); print('$wl value %.4f e2e %.4f ms %.2f frac %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']), d['config'].get('decoder_path'))
except Exception as e:
    print('$wl failed', e); print(open('gpurun_out/r2c_$wl.err').read()[-2000:])
"
done
for t in 384 512 640 768; do
  python bench.py --workload I80_nmsa_q030 --onchip-threads $t --frames 16384 --steps 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r2c_t$t.json 2> gpurun_out/r2c_t$t.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2c_t$t.json')); print('threads $t value %.4f'%d['value'])
except Exception as e: print('threads $t failed', e)
"
done
