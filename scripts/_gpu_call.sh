mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py tests/test_gpu_training.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|Error|assert|^E " | head -60 > gpurun_out/t.log
cat gpurun_out/t.log
