timeout 900 python -m pytest tests/test_gpu_occmap_tc.py -m gpu -x -q -s 2>&1 | grep -E "render_rays|passed|failed|^E  " > gpurun_out/t.log
cat gpurun_out/t.log
