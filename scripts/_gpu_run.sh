V=category_nerf_reconstruction_official_b200/csrc/variants
timeout 900 python -m pytest tests/test_gpu_parity_v0.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t.log
for r in 1 2; do
CNF_LIB_PATH=$V/libcnf_smp1.so timeout 300 python scripts/bench_standalone.py 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('old', d['sampler'])" >> gpurun_out/t.log
timeout 300 python scripts/bench_standalone.py 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('new', d['sampler'])" >> gpurun_out/t.log
done
cat gpurun_out/t.log
