mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wrappers.py -m gpu -x -q -k pipeline 2>&1 | tail -30 > gpurun_out/tests.log
cat gpurun_out/tests.log
