timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "rc=$?" > gpurun_out/t.log
python -c "import json; d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e']['value'], sorted(k for k in d if k.endswith('cube') or 'step' in k or 'c4' in k))" >> gpurun_out/t.log 2>&1
tail -3 gpurun_out/bench2.err >> gpurun_out/t.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/ref2.json 2>/dev/null; echo "ref rc=$? lines=$(wc -l < gpurun_out/ref2.json)" >> gpurun_out/t.log
cat gpurun_out/t.log
