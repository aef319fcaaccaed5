timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/t.log 2>&1
timeout 300 python bench.py --no-large 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])" >> gpurun_out/t.log
cat gpurun_out/t.log
