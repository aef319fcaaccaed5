timeout 900 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q -s 2>&1 | grep -E "^P=|background|passed|failed|^E  " | head -30 > gpurun_out/t.log
cat gpurun_out/t.log
