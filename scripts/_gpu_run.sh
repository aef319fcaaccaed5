timeout 600 python -m pytest tests/test_gpu_parity_v0.py -m gpu -x -q -k origin_dirs 2>&1 | tail -3 > gpurun_out/t.log
timeout 300 python scripts/profile_bg.py 3 >> gpurun_out/t.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^bwd_kernel" --launch-skip 2 --launch-count 1 -f -o gpurun_out/otc_bwd python scripts/profile_bg.py 3 > gpurun_out/ncu_otc.log 2>&1
tail -3 gpurun_out/ncu_otc.log >> gpurun_out/t.log
cat gpurun_out/t.log
