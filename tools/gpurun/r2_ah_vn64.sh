# round 2, call AH: walking variable node for float64 at 3 CTAs per SM (77 registers) and for 2 frames per lane at 5 / 6 CTAs
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "1:0 8:3 29:3 8:1" 1 0 64 2>&1 | tee gpurun_out/r2ah_l100k_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 8192 0 0 0.0162 "1:0 8:3 4:3" 2 0 64 2>&1 | tee gpurun_out/r2ah_a82_spa64.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 32768 0 0 0.0162 "1:0 8:5 8:6 4:5" 2 0 32 2>&1 | tee gpurun_out/r2ah_a82_spa32.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A82 8192 5 0.68 0.0161 "1:0 8:3" 2 0 64 2>&1 | tee gpurun_out/r2ah_a82_aomsa64.txt | grep -v Warning
