timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 5 --no-large > gpurun_out/bench8.json 2> gpurun_out/bench8.err
python -c "import json; d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['e2e'])" > gpurun_out/t.log 2>&1
tail -2 gpurun_out/bench8.err >> gpurun_out/t.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 10 --warmup 3 --ray-split 2>/dev/null | tail -1 | cut -c1-400 >> gpurun_out/t.log
cat gpurun_out/t.log
