for i in 1 2; do timeout 300 python bench.py --no-large --no-cpu-baseline 2>gpurun_out/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e'])"; done > gpurun_out/t.log 2>&1
tail -3 gpurun_out/bench.err >> gpurun_out/t.log
cat gpurun_out/t.log
