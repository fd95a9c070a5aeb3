# round 2, call AY: walking kernel for the wide variable-node buckets (8 < dv <= 32) against vn_kernel (vn_items_per_warp = 1)
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_onchip.py tests/test_gpu_random_codes.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.03 "1:0 0:0" 2 2>&1 | grep -v Warning | tee gpurun_out/r2ay_wide.txt
timeout 120 python tools/vn_sweep.py I80 32768 0 0 0.015 "1:0 0:0" 2 0 32 2>&1 | grep -v Warning | tee -a gpurun_out/r2ay_wide.txt
timeout 120 python tools/vn_sweep.py I80 8192 2 0.7 0.015 "1:0 0:0" 2 0 64 2>&1 | grep -v Warning | tee -a gpurun_out/r2ay_wide.txt
