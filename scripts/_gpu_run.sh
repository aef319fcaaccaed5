timeout 600 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q -k capture_iteration 2>&1 | tail -15 > gpurun_out/t.log
cat gpurun_out/t.log
