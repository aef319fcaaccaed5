timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^E  |well-cond" | head -20 > gpurun_out/t.log
cat gpurun_out/t.log
