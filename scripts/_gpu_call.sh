mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -E "^\[|passed|failed|Error|assert|max abs|rel|P=|background" > gpurun_out/t.log
tail -30 gpurun_out/t.log
timeout 600 python scripts/seed_sweep.py > gpurun_out/seed_sweep.txt 2>&1; tail -3 gpurun_out/seed_sweep.txt
