timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -2 gpurun_out/bench.err >> gpurun_out/t.log
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
for k in ('large_batch_c4','background_step_1200x14','iteration_objects_plus_background'): print(k, {kk: (round(v, 4) if isinstance(v, float) else v) for kk, v in d.get(k, {}).items() if kk not in ('note','config')})
" >> gpurun_out/t.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/t.log 2>&1
cat gpurun_out/t.log
