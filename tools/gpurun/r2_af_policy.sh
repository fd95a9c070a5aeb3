# round 2, call AF: compaction only when the batch is expected to last (retire rate), default fill 75 %; the streaming tests
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_onchip.py tests/test_gpu_random_codes.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "0:0:16 0:0 0:0:0:0:1" 3 2>&1 | tee gpurun_out/r2af_l100k.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "0:0:16:50 0:0 0:0:4" 2 2>&1 | tee gpurun_out/r2af_a79.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 4096 2 0.71 0.02 "0:0:16:50 0:0 0:0:16" 3 2>&1 | tee gpurun_out/r2af_a79_small.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.015 "0:0:16:50 0:0" 2 2>&1 | tee gpurun_out/r2af_i80_q015.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "0:0:16:50 0:0" 1 0 64 2>&1 | tee gpurun_out/r2af_l100k_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 8192 0 0 0.0162 "0:0:16:50 0:0" 2 0 64 2>&1 | tee gpurun_out/r2af_a82_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 32768 0 0 0.0162 "0:0:16:50 0:0" 2 0 32 2>&1 | tee gpurun_out/r2af_a82_spa32.txt | grep -v Warning
