mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hbm_kernels.py tests/test_gpu_parity_v0.py tests/test_gpu_occmap_tc.py tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/tests.log
cat gpurun_out/tests.log
timeout 600 python scripts/bench_standalone.py > gpurun_out/standalone.json 2> gpurun_out/standalone.err; cat gpurun_out/standalone.json; tail -3 gpurun_out/standalone.err
CNF_PROFILE_ONCE=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r02_hbm_ncu.csv python scripts/bench_standalone.py > gpurun_out/r02_hbm_ncu.log 2>&1
python scripts/hbm_ncu_table.py gpurun_out/r02_hbm_ncu.csv
