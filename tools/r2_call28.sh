#!/bin/bash
# round 2, GPU call 28 (2 GPUs): pipelined all-reduce + AdamW under NCCL with graphs -- 2-rank parity test, then 2-rank bench off / on
set -x
O=gpurun_out/r2c28
mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_dist_gpu.py -q -m gpu --tb=short -rA -p no:cacheprovider > $O/dist_test.log 2>&1; echo "dist test rc=$?" | tee $O/rc.txt
tail -8 $O/dist_test.log
for mode in 0 1 0 1; do
  T0=$(date +%s)
  TVTS_PIPELINED_ADAMW=$mode timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29512 + RANDOM % 100)) bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_2gpu_pipe$mode.json 2> $O/bench_2gpu_pipe$mode.err
  echo "2gpu pipelined=$mode rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; python -c "
import json,sys
d=json.loads(open('$O/bench_2gpu_pipe$mode.json').read().strip().splitlines()[-1]); print('pipelined=$mode', round(d['value'],1), round(d['ms_per_step'],3), d['config']['step'])"
done
