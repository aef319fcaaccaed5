timeout 300 python scripts/time_bg_eval.py plain > gpurun_out/t.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^eval_kernel" --launch-skip 10 --launch-count 1 -f -o gpurun_out/otc_eval python scripts/time_bg_eval.py ncu > gpurun_out/ncu_eval.log 2>&1
tail -2 gpurun_out/ncu_eval.log >> gpurun_out/t.log
cat gpurun_out/t.log
