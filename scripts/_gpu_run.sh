timeout 900 python -m pytest tests/test_gpu_occmap_tc.py tests/test_gpu_parity_v0.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
timeout 300 python scripts/time_bg.py tc 2>&1 | grep -E "Fused" >> gpurun_out/t.log
cat gpurun_out/t.log
