mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_properties.py tests/test_gpu_hbm_kernels.py tests/test_gpu_parity_v0.py -m gpu -q 2>&1 | grep -E "Falsifying|seed=|K=|R=|N=|M=|S=|P=|n=|inv=|p_this=|Error|assert|passed|failed" | head -40 > gpurun_out/tests.log
cat gpurun_out/tests.log
