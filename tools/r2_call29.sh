#!/bin/bash
# round 2, GPU call 29 (8 GPUs): the final code at 8 ranks -- c3 bench with the single all-reduce (default) and with the pipelined all-reduce + AdamW
set -x
O=gpurun_out/r2c29
mkdir -p $O
nvidia-smi -L | wc -l > $O/gpus.txt
for mode in 0 1; do
  T0=$(date +%s)
  TVTS_PIPELINED_ADAMW=$mode timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > $O/bench_8gpu_pipe$mode.json 2> $O/bench_8gpu_pipe$mode.err
  echo "8gpu pipelined=$mode rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt
  tail -c 300 $O/bench_8gpu_pipe$mode.json | head -c 300; echo
done
CUDA_VISIBLE_DEVICES=0 timeout -k 10 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-e2e > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 200 $O/bench_1gpu.json
