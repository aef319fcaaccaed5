mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests.log
cat gpurun_out/tests.log
for tool in memcheck synccheck racecheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log
  tail -5 gpurun_out/sanitizer_$tool.log
done
