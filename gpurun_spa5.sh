timeout 1200 python -m pytest tests/test_gpu_large.py -x -q -s -k "A82-0" 2>&1 | grep -E "iterations equal|passed|failed"
