V=category_nerf_reconstruction_official_b200/csrc/variants
timeout 600 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q -s -k eval_points 2>&1 | grep -E "eval_points|passed|failed|^E  " > gpurun_out/t.log
for r in 1 2; do
CNF_LIB_PATH=$V/libcnf_ev1cta.so timeout 300 python scripts/time_bg_eval.py 1cta >> gpurun_out/t.log 2>&1
timeout 300 python scripts/time_bg_eval.py 2cta >> gpurun_out/t.log 2>&1
done
cat gpurun_out/t.log
