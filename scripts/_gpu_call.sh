mkdir -p gpurun_out
CNF_HYPOTHESIS=explore timeout 2400 python -m pytest tests/test_properties.py -m gpu -q 2>&1 | grep -E "Falsifying|seed=|K=|n_obj=|R=|N=|M=|L=|S=|P=|n=|bins=|scale=|inv=|p_this|Error|assert|passed|failed|Mismatch|Greatest" | head -60 > gpurun_out/tests.log
cat gpurun_out/tests.log
timeout 2400 python -m pytest tests/test_properties.py -m gpu -q 2>&1 | tail -2
