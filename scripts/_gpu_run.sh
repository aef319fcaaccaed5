timeout 120 python scripts/check_otc.py > gpurun_out/t.log 2>&1
cat gpurun_out/t.log
