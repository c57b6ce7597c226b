#!/bin/bash
# round 2, GPU call 26 (2 GPUs): the final code under NCCL -- 2-rank parity test (graphs, different caption lengths per rank, clean teardown), 2-rank bench, 1-rank bench
set -x
O=gpurun_out/r2c26
mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_dist_gpu.py -q -m gpu --tb=short -rA -p no:cacheprovider > $O/dist_test.log 2>&1; echo "dist test rc=$?" | tee $O/rc.txt
tail -5 $O/dist_test.log
T0=$(date +%s)
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "2gpu rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -c 400 $O/bench_2gpu.json
CUDA_VISIBLE_DEVICES=0 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.json
