for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json')); print('N=$n value %.3f Gbit/s e2e %.3f ms/step %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['clocks'])" || tail -5 gpurun_out/bench_n$n.err
done
