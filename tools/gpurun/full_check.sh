timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; tail -c 1500 gpurun_out/bench_r01f.json; tail -3 gpurun_out/bench_r01f.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r01f_ref.json 2>> gpurun_out/bench_r01f.err; cat gpurun_out/bench_r01f_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
