V=category_nerf_reconstruction_official_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q 2>&1 | grep -E "passed|failed|^E  |FAILED" | head -5 > gpurun_out/t.log
for r in 1 2; do
CNF_LIB_PATH=$V/libcnf_bwd160.so timeout 300 python scripts/time_otc.py old 2>&1 | head -1 >> gpurun_out/t.log
timeout 300 python scripts/time_otc.py new 2>&1 | head -1 >> gpurun_out/t.log
done
timeout 300 python scripts/time_bg.py tc 2>&1 | grep -E "graph" >> gpurun_out/t.log
cat gpurun_out/t.log
