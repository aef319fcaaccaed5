timeout 300 python scripts/finalize_timing.py 2>&1 | sed -n '/prepare/,$p' > gpurun_out/t.log
cat gpurun_out/t.log
