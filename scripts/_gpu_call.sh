mkdir -p gpurun_out
CNF_HYPOTHESIS=explore timeout 2400 python -m pytest tests/test_properties.py -m gpu -q -k fused 2>&1 | grep -E "Falsifying|seed=|K=|n_obj=|R=|N=|M=|L=|S=|P=|n=|inv=|p_this|Error|assert|passed|failed|rel L2" | head -60 > gpurun_out/tests.log
cat gpurun_out/tests.log
