# round 2, call AD: the walking kernel under its auto plan; steps between two host polls and the occupancy at which the tail is compacted
export VN_SWEEP_HIST=1
timeout 200 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "1:0 0:0 0:0:8 0:0:4 0:0:2 0:0:1 0:0:2:75 0:0:1:75 0:0:1:90 0:0:4:75" 3 2>&1 | tee gpurun_out/r2ad_l100k.txt | grep -v Warning
unset VN_SWEEP_HIST
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.03 "1:0 0:0" 2 2>&1 | tee gpurun_out/r2ad_i80.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py I80 32768 2 0.7 0.015 "1:0 0:0 0:0:4 0:0:4:75 0:0:2:75" 2 2>&1 | tee gpurun_out/r2ad_i80_q015.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "1:0 0:0 0:0:4 0:0:4:75 0:0:2:75" 2 2>&1 | tee gpurun_out/r2ad_a79.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py A79 4096 2 0.71 0.02 "1:0 0:0 0:0:4 0:0:4:75 0:0:2:75 0:0:1:75" 3 2>&1 | tee gpurun_out/r2ad_a79_small.txt | grep -v Warning
timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "1:0 0:0:4 0:0:4:75 0:0:2:75" 1 0 64 2>&1 | tee gpurun_out/r2ad_l100k_spa64.txt | grep -v Warning
