python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 900 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 1 --warmup 1 --impl reference > gpurun_out/bench_n2_ref.json 2>> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2_ref.json | cut -c1-300
