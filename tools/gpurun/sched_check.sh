# scheduler kernels after the retire rewrite: streaming parity tests, n = 102400 throughput, ncu launch list of that command
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_random_codes.py tests/test_sim_csv.py -x -q 2>&1 | tail -2
CMD="python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD 2> gpurun_out/plain_l100k.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('L100k %.3f Gbit/s sched %.2f ms/step whole %.3f'%(d['value'], d['roofline']['sched_ms_per_step'], d['roofline']['whole_step_frac']), {k:round(v['frac'],3) for k,v in d['roofline']['both_kernels'].items()})"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_l100k.csv python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ll_l100k.log 2>&1
