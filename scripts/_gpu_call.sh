mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests.log
cat gpurun_out/tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','cpu_baseline')}); print(d['e2e']); print(d['roofline'])
PY
