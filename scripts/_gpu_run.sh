timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
timeout 600 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -2 gpurun_out/bench.err >> gpurun_out/t.log
python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
for k in ('large_batch_c4','eval_points_256cube','occmap_eval_points_256cube','render_512x512_S14','background_step_1200x14','iteration_objects_plus_background'): print(k, {kk: (round(v, 4) if isinstance(v, float) else v) for kk, v in d.get(k, {}).items() if kk != 'note' and kk != 'config'})
" >> gpurun_out/t.log 2>&1
timeout 300 python scripts/run_fused.py replica_room0 360 6 >> gpurun_out/t.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2_launches.csv python scripts/run_fused.py replica_room0 360 6 > gpurun_out/ncu_c2.log 2>&1
cat gpurun_out/t.log
