mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity_r2.py -m gpu -q -s -k "curve" 2>&1 | grep -E "loss-curve|loss curve|passed|failed|early|diverge" ; done > gpurun_out/t.log
cat gpurun_out/t.log
