# round 2, call AC: vn_kernel_ell_loop without spills (3 / 4 CTAs per SM for dv <= 4, 2 for dv <= 8), longer walks, float64
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vn_items" 2>&1 | tail -3
timeout 200 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "1:0 1:4 8:4 16:4 32:4 64:4 8:3 16:3 32:3 64:3" 3 2>&1 | tee gpurun_out/r2ac_l100k.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.03 "1:0 8:4 16:4 16:3 16:6" 2 2>&1 | tee gpurun_out/r2ac_i80.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "1:0 8:4 16:4 32:4 16:3" 2 2>&1 | tee gpurun_out/r2ac_a79.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "1:0 8:4 32:4" 1 0 64 2>&1 | tee gpurun_out/r2ac_l100k_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 8192 0 0 0.0162 "1:0 8:4 32:4" 2 0 64 2>&1 | tee gpurun_out/r2ac_a82_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 32768 0 0 0.0162 "1:0 8:4 32:4" 2 0 32 2>&1 | tee gpurun_out/r2ac_a82_spa32.txt | grep -v Warning
