timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t.log
timeout 300 python scripts/bench_standalone.py > gpurun_out/sa.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/sa.json')); print({k: (round(v['ms'],4), round(v['gbs']), round(v['frac'],3)) for k,v in d.items() if isinstance(v, dict)})" >> gpurun_out/t.log
timeout 300 python scripts/race_check.py 2>&1 | tail -3 >> gpurun_out/t.log
timeout 300 python scripts/garbage_check.py 2>&1 | awk '{print $NF, $(NF-3)}' | sort | uniq -c | head -5 >> gpurun_out/t.log
cat gpurun_out/t.log
