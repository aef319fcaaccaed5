timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t.log
timeout 300 python scripts/race_check.py 2>&1 | tail -3 >> gpurun_out/t.log
for i in 1 2; do
timeout 300 python bench.py --no-large --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'eager', d['config']['eager_ms_per_step'])" >> gpurun_out/t.log
done
cat gpurun_out/t.log
