# round 2, call AE: step-graph cache, poll interval from the step time, quantised tile counts, parallel compaction plan
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_onchip.py -m gpu -x -q -k "compaction or small_pool or vn_items or refill or graph" 2>&1 | tail -4
timeout 200 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "1:0:16 0:0:16 0:0 0:0:1 0:0:2 0:0:4 0:0:0:0:1 0:0:0:75" 3 2>&1 | tee gpurun_out/r2ae_l100k.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "0:0:16 0:0 0:0:0:75 0:0:4 0:0:4:75 0:0:1 0:0:0:0:1" 2 2>&1 | tee gpurun_out/r2ae_a79.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 4096 2 0.71 0.02 "0:0:16 0:0 0:0:4 0:0:2 0:0:0:75" 3 2>&1 | tee gpurun_out/r2ae_a79_small.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.015 "0:0:16 0:0 0:0:0:75 0:0:4:75" 2 2>&1 | tee gpurun_out/r2ae_i80_q015.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "0:0:16 0:0 0:0:0:75 0:0:0:0:1" 1 0 64 2>&1 | tee gpurun_out/r2ae_l100k_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 8192 0 0 0.0162 "0:0:16 0:0 0:0:0:75" 2 0 64 2>&1 | tee gpurun_out/r2ae_a82_spa64.txt | grep -v Warning
