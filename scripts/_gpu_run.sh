timeout 900 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q -s 2>&1 | grep -E "points variant|background trainer|passed|failed|^E  " | head -20 > gpurun_out/t.log
timeout 300 python scripts/time_bg.py tc 2>&1 | grep -E "Fused" >> gpurun_out/t.log
cat gpurun_out/t.log
