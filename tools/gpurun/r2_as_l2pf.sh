# round 2, call AS: walking variable node with an L2 prefetch of the next item's message chunks (QKDLDPC_VN_L2PF=1) against without
for pf in 0 1; do
  echo "== QKDLDPC_VN_L2PF=$pf"
  export QKDLDPC_VN_L2PF=$pf
  timeout 120 python tools/vn_sweep.py L100k 1024 0 0 0.084 "0:0" 2 0 64 2>&1 | grep -v Warning | tee -a gpurun_out/r2as_l2pf.txt
  timeout 120 python tools/vn_sweep.py L100k 4096 2 0.72 0.06 "0:0" 3 2>&1 | grep -v Warning | tee -a gpurun_out/r2as_l2pf.txt
  timeout 120 python tools/vn_sweep.py A82 8192 0 0 0.0162 "0:0" 2 0 64 2>&1 | grep -v Warning | tee -a gpurun_out/r2as_l2pf.txt
  timeout 120 python tools/vn_sweep.py A79 32768 2 0.71 0.02 "0:0" 2 2>&1 | grep -v Warning | tee -a gpurun_out/r2as_l2pf.txt
done
export QKDLDPC_VN_L2PF=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vn_items or compaction or small_pool" 2>&1 | tail -2
