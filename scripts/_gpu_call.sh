mkdir -p gpurun_out
CNF_HYPOTHESIS=explore timeout 2400 python -m pytest tests/test_properties.py -m gpu -q -k "embedding or termination or stratified" 2>&1 | grep -E "Falsifying|seed=|K=|P=|n=|S=|bins=|scale=|Error|assert|passed|failed|Mismatch|Greatest" | head -60 > gpurun_out/tests.log
cat gpurun_out/tests.log
