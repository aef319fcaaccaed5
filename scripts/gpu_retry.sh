#!/bin/bash
# development aid: run scripts/_gpu_call.sh on a GPU box, retrying while the pod is busy.  usage: scripts/gpu_retry.sh [gpus] [timeout]
G=${1:-1}; T=${2:-1800}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- 'bash scripts/_gpu_call.sh' > gpurun_out/call.log 2>&1
  else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- 'bash scripts/_gpu_call.sh' > gpurun_out/call.log 2>&1; fi
  if grep -q "status=transient" gpurun_out/call.log; then sleep 60; continue; fi
  break
done
tail -100 gpurun_out/call.log
