mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_hbm_kernels.py tests/test_gpu_parity_v0.py -m gpu -q 2>&1 | grep -E "Mismatch|Greatest|passed|failed" | head > gpurun_out/tests.log
cat gpurun_out/tests.log
timeout 600 python scripts/bench_standalone.py > gpurun_out/standalone.json 2> gpurun_out/standalone.err; cat gpurun_out/standalone.json | cut -c 1-900; tail -3 gpurun_out/standalone.err
