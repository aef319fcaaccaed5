timeout 300 python scripts/finalize_timing.py 2>&1 | grep -E "finalize:|job 4 lat s0 p0|job 5 codes p0|job 5 codes p3" > gpurun_out/t.log
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_training.py tests/test_gpu_parity_v0.py -m gpu -x -q 2>&1 | tail -1 >> gpurun_out/t.log
timeout 300 python scripts/race_check.py 2>&1 | tail -3 >> gpurun_out/t.log
for i in 1 2; do
timeout 300 python bench.py --no-large --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" >> gpurun_out/t.log
done
cat gpurun_out/t.log
