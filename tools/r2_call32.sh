#!/bin/bash
# round 2, GPU call 32 (4 GPUs): the final code at 1, 2 and 4 ranks of one box (the 8-rank line is call 29's)
set -x
O=gpurun_out/r2c32
mkdir -p $O
CUDA_VISIBLE_DEVICES=0 timeout -k 10 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-e2e > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "1gpu rc=$?" | tee $O/rc.txt
for n in 2 4; do
  timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29300 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
  echo "n=$n rc=$?" | tee -a $O/rc.txt
done
python - <<'PY'
import json
for n in (1, 2, 4):
    d = json.loads(open(f"gpurun_out/r2c32/bench_{n}gpu.json").read().strip().splitlines()[-1])
    print(n, round(d["value"], 1), round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
PY
