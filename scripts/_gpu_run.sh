timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t.log
timeout 300 python scripts/race_check.py 2>&1 | tail -5 >> gpurun_out/t.log
timeout 300 python bench.py --steps 50 --warmup 5 2>&1 | tail -2 >> gpurun_out/t.log
cat gpurun_out/t.log
