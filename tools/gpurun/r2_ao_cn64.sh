# round 2, call AO: float64 sum-product check node with the row's tanh values in registers (dc <= 8 bucket) instead of the two-pass kernel
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "0:0" 2 0 64 2>&1 | tee gpurun_out/r2ao_l100k_spa64.txt | grep -v Warning
timeout 300 python -m pytest tests/test_gpu_points.py tests/test_gpu_parity.py -m gpu -x -q -k "spa or 100k or fp64" 2>&1 | tail -3
