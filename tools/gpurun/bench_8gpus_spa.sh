# SPA workload (A82 @ QBER 1.62 %) on 8, 4, 2 and 1 GPUs of one box: weak scaling, tally all-reduce only
for n in 8 4 2 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload A82_spa_q0162 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_spa_n$n.json 2> gpurun_out/bench_spa_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_spa_n$n.json')); print('SPA N=$n value %.3f Gbit/s e2e %.3f ms/step %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'])" || tail -5 gpurun_out/bench_spa_n$n.err
done
