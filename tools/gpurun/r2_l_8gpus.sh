# round 2, call L (8 GPUs): strong scaling of the headline workload (262 144 frames per step over all GPUs), the default weak-
# scaling line with its workloads at N = 8, and whole configs through qkdldpc_sim --gpus 8 (NCCL tally all-reduce in the library)
for n in 1 2 4 8; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n"; fi
  $L bench.py --gpus $n --steps 2 --warmup 1 --scaling strong --total-frames 262144 --no-cpu-baseline --no-secondary > gpurun_out/r2l_strong_n$n.json 2> gpurun_out/r2l_strong_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/r2l_strong_n$n.json')); print('strong N=$n value %.4f Gbit/s e2e %.4f ms/step %.1f frames/gpu %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['config'].get('frames_per_step_per_gpu')))" || tail -5 gpurun_out/r2l_strong_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29629 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_weak_n8.json 2> gpurun_out/r2l_weak_n8.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2l_weak_n8.json'))
    print('weak N=8 HEAD value %.4f e2e %.4f' % (d['value'], d['e2e']['value']))
    for w in d['workloads']:
        print('  %-18s value %.3f e2e %.3f dtype %s it %.2f fer %.4f %s' % (w['workload_id'], w['value'], w['e2e']['value'], w['dtype'], w['mean_iterations_executed'], w['fer'], w['decoder_path'][:12]))
except Exception as e:
    print('weak N=8 failed', e); print(open('gpurun_out/r2l_weak_n8.err').read()[-2000:])
PY
for cfg in config10k config100k config100k_nmsa; do
  python tools/config_parity.py --config $cfg --ref-trials 1000 --gpus 8 --full --out gpurun_out/r2l_$cfg.json > /dev/null 2> gpurun_out/r2l_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2l_$cfg.json')); r=d['runs']
    print('$cfg: ref cpu %.1fs (%d trials)'%(r['reference_cpu']['seconds'], d['ref_trials']))
    for k in ('qkdldpc_sim_default','qkdldpc_sim_fp32','qkdldpc_sim_fp64'):
        rows=r[k]['rows']
        print('  ',k,'%.1fs'%r[k]['seconds'],'rows',len(rows),'csv identical',r[k]['csv_identical'],'fer inside ci',sum(x['fer_inside_ci'] for x in rows))
    f=r['qkdldpc_sim_full']; print('   full on 8 GPUs: %.1fs for %d combinations x %d trials; reference extrapolated %.0fs'%(f['seconds'],f['combinations'],f['trials'],f['reference_cpu_seconds_extrapolated']))
except Exception as e:
    print('$cfg failed', e); print(open('gpurun_out/r2l_$cfg.err').read()[-1500:])
PY
done
