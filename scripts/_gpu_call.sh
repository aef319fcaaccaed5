mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_occmap_tc.py tests/test_gpu_fused.py -m gpu -q -s 2>&1 | grep -E "^\[|P=|background|eval_points|render_rays|well-cond|passed|failed|Error|assert" > gpurun_out/t.log
cat gpurun_out/t.log
