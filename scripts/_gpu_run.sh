timeout 900 python -m pytest tests/test_gpu_parity_v0.py -m gpu -x -q 2>&1 | tail -1 > gpurun_out/t.log
timeout 300 python scripts/bench_standalone.py > gpurun_out/sa.json 2> gpurun_out/sa.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sample_" -c 6 --csv --log-file gpurun_out/sampler_launches.csv python scripts/bench_standalone.py > gpurun_out/ncu_sa.log 2>&1
python -c "import json; d=json.load(open('gpurun_out/sa.json')); print(d['sampler'])" >> gpurun_out/t.log
grep -E "sample_" gpurun_out/sampler_launches.csv | awk -F'","' '{print substr($5,1,24), $(NF)}' | tail -3 >> gpurun_out/t.log
cat gpurun_out/t.log
