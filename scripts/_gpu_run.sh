timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err >> gpurun_out/t.log
timeout 300 python scripts/time_bg.py tc 2>&1 | grep -E "Fused" >> gpurun_out/t.log
cat gpurun_out/t.log; cat gpurun_out/bench.json
