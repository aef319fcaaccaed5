V=category_nerf_reconstruction_official_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "render_rays or eval_points" 2>&1 | tail -3 > gpurun_out/t.log
for r in 1 2; do
CNF_LIB_PATH=$V/libcnf_ev3.so timeout 300 python scripts/time_render.py 2>&1 | sed 's/^/3 chains: /' >> gpurun_out/t.log
timeout 300 python scripts/time_render.py 2>&1 | sed 's/^/4 chains: /' >> gpurun_out/t.log
CNF_LIB_PATH=$V/libcnf_ev3.so timeout 300 python scripts/time_eval_points.py 2>&1 | tail -2 | sed 's/^/3 chains: /' >> gpurun_out/t.log
timeout 300 python scripts/time_eval_points.py 2>&1 | tail -2 | sed 's/^/4 chains: /' >> gpurun_out/t.log
done
cat gpurun_out/t.log
