mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_fused.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests.log
cat gpurun_out/tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv python scripts/run_fused.py c4 524288 4 > gpurun_out/l2.log 2>&1
python scripts/launch_list.py gpurun_out/r02_launches_c4.csv
