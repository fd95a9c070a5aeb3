# the north-star parity triple on the reference configs' own trial counts (10 000 - 30 000 frames per operating point)
QKD_PARITY_FULL=1 timeout 2400 python -m pytest tests/test_gpu_large.py -x -q -s 2>&1 | grep -E "alg=|passed|failed|Error" | tee gpurun_out/parity_full.txt
