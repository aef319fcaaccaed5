mkdir -p gpurun_out
bash scripts/ab.sh base sleep20 sleep60 nohint 2>&1 | grep -v Warning > gpurun_out/ab.txt; cat gpurun_out/ab.txt
