timeout 300 python scripts/run_fused.py replica_room0 360 4 > gpurun_out/t.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_finalize --launch-skip 2 --launch-count 1 -f -o gpurun_out/fin2 python scripts/run_fused.py replica_room0 360 4 > gpurun_out/ncu_fin.log 2>&1
cat gpurun_out/t.log
