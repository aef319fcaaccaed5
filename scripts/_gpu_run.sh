timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err > gpurun_out/t.log
python -c "import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d.get('background_step_1200x14'), d.get('iteration_objects_plus_background'))" >> gpurun_out/t.log 2>&1
cat gpurun_out/t.log
