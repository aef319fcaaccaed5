# round 2, call A: baseline GPU tests + ncu DRAM counters of the standalone HBM-bound kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2a_tests.log
timeout 600 python scripts/bench_standalone.py > gpurun_out/r2a_standalone.json 2> gpurun_out/r2a_standalone.err
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 800 --csv \
  --log-file gpurun_out/r2a_hbm_ncu.csv python scripts/bench_standalone.py > gpurun_out/r2a_hbm_ncu.log 2>&1
cat gpurun_out/r2a_tests.log
