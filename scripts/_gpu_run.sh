timeout 900 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q -s 2>&1 | grep -E "background|passed|failed|^E  |Error" | head -30 > gpurun_out/t.log
timeout 300 python scripts/time_bg.py tc 2>&1 | grep -E "bg step|Fused|Error|error" >> gpurun_out/t.log
cat gpurun_out/t.log
