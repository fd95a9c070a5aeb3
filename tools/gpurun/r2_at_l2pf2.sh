# round 2, call AT: L2 prefetch distance 1 / 2 / 3 items in the walking variable node
for pf in 1 2 3; do
  echo "== QKDLDPC_VN_L2PF=$pf"
  export QKDLDPC_VN_L2PF=$pf
  timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "0:0" 2 0 64 2>&1 | grep -v Warning | tee -a gpurun_out/r2at_l2pf.txt
  timeout 120 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "0:0" 3 2>&1 | grep -v Warning | tee -a gpurun_out/r2at_l2pf.txt
  timeout 120 python tools/vn_sweep.py A82 8192 0 0 0.0162 "0:0" 2 0 64 2>&1 | grep -v Warning | tee -a gpurun_out/r2at_l2pf.txt
  timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "0:0" 2 2>&1 | grep -v Warning | tee -a gpurun_out/r2at_l2pf.txt
done
