timeout 300 python scripts/profile_bg.py 3 > gpurun_out/t.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/bg_launches.csv python scripts/profile_bg.py 3 > gpurun_out/ncu_bg.log 2>&1
cat gpurun_out/t.log
