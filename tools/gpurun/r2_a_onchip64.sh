# round 2, call A: the float64 on-chip min-sum kernel -- GPU test-suite, then bench lines of the offset / adaptive workloads
# under the default precision policy (float64 state) and with float32 forced, and the headline in float64 for comparison.
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt
cat gpurun_out/r2a_pytest.txt
b() {  # name, extra args...
  name=$1; shift
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2a_$name.json 2> gpurun_out/r2a_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2a_$name.json')); print('$name value %.3f e2e %.3f dtype %s path %s it %.2f fer %.4f frac %.2f'%(d['value'], d['e2e']['value'], d['dtype'], d['config']['decoder_path'], d['config']['mean_iterations_executed'], d['config']['fer'], d['roofline']['frac']))" || tail -5 gpurun_out/r2a_$name.err
}
b A82_aomsa_auto --workload A82_aomsa_q0161
b A82_aomsa_f32 --workload A82_aomsa_q0161 --precision 32
b A82_aomsa_f64_streaming --workload A82_aomsa_q0161 --path 1
b A82_omsa_auto --workload A82_omsa_q0154
b A82_anmsa_auto --workload A82_anmsa_q0161
b A82_anmsa_f32 --workload A82_anmsa_q0161 --precision 32
b I80_aomsa_auto --workload I80_aomsa_q015
b I80_nmsa_q030_f64 --workload I80_nmsa_q030 --precision 64
b I80_nmsa_q030_f32 --workload I80_nmsa_q030
