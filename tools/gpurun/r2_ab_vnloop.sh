# round 2, call AB: vn_kernel_ell_loop (several items per warp, L1 prefetch of the next item's index records) against vn_kernel_ell
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vn_items or tail_compaction or small_pool" 2>&1 | tail -4
S="1:0 2:0 4:0 8:0 16:0 4:5 8:5 4:4 8:4"
timeout 200 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "$S" 3 2>&1 | tee gpurun_out/r2ab_l100k.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py L100k 16384 2 0.72 0.06 "1:0 4:0 4:5" 2 4096 2>&1 | tee gpurun_out/r2ab_l100k_refill.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.03 "1:0 4:0 8:0 4:5 4:4" 2 2>&1 | tee gpurun_out/r2ab_i80.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "1:0 4:0 8:0 4:5 4:4" 2 2>&1 | tee gpurun_out/r2ab_a79.txt | grep -v Warning
