timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-large > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -2 gpurun_out/bench2.err > gpurun_out/t.log
python -c "import json; d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e'])" >> gpurun_out/t.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300 >> gpurun_out/t.log
cat gpurun_out/t.log
