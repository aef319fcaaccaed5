timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_training.py tests/test_gpu_parity_v0.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
timeout 300 python scripts/race_check.py 2>&1 | tail -3 >> gpurun_out/t.log
timeout 300 python scripts/run_fused.py c4 524288 4 >> gpurun_out/t.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c4_launches.csv python scripts/run_fused.py c4 524288 4 > gpurun_out/ncu_c4.log 2>&1
cat gpurun_out/t.log
