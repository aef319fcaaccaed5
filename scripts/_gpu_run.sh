timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err >> gpurun_out/t.log
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
for k in ('large_batch_c4','eval_points_256cube','render_512x512_S14','render_512x512_S96','background_step_1200x14','iteration_objects_plus_background'): print(k, d.get(k))
" >> gpurun_out/t.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> gpurun_out/t.log 2>&1
cat gpurun_out/t.log
