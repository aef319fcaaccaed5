timeout 900 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
timeout 300 python scripts/time_otc.py pts >> gpurun_out/t.log 2>&1
timeout 300 python scripts/stamps_otc.py 16800 pts 2>&1 | grep -E "backward|S0|S5" >> gpurun_out/t.log
timeout 300 python scripts/time_bg.py tc 2>&1 | grep -E "Fused" >> gpurun_out/t.log
cat gpurun_out/t.log
