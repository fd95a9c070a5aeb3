# round 2, call AK: lanes whose active slots were all just refilled skip the message loads (first step of a batch)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_codes.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "0:0 1:0" 3 2>&1 | tee gpurun_out/r2ak_l100k.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "0:0" 1 0 64 2>&1 | tee gpurun_out/r2ak_l100k_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "0:0" 2 2>&1 | tee gpurun_out/r2ak_a79.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 32768 0 0 0.0162 "0:0" 2 0 32 2>&1 | tee gpurun_out/r2ak_a82_spa32.txt | grep -v Warning
