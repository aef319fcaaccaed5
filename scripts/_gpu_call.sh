mkdir -p gpurun_out
timeout 600 python scripts/bench_standalone.py > gpurun_out/standalone.json 2> gpurun_out/standalone.err; cat gpurun_out/standalone.json; tail -3 gpurun_out/standalone.err
CNF_PROFILE_ONCE=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r02_hbm_ncu.csv python scripts/bench_standalone.py > gpurun_out/r02_hbm_ncu.log 2>&1
python scripts/hbm_ncu_table.py gpurun_out/r02_hbm_ncu.csv
timeout 600 python scripts/profile_bg.py 3 > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_background_step.csv python scripts/profile_bg.py 3 > gpurun_out/l4.log 2>&1
timeout 1200 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -2 gpurun_out/bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['cpu_baseline']['ms_per_step'], d['large_batch_c4'])
PY
