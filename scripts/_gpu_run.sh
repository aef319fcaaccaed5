timeout 300 python scripts/stamps_otc.py 16800 > gpurun_out/stamps_otc.txt 2>&1
cat gpurun_out/stamps_otc.txt | tail -45
