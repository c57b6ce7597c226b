#!/bin/bash
# round 2, GPU call 6 (2 GPUs): 2-rank NCCL parity test (graphs + bucketed overlap + different caption lengths per rank), clean teardown,
# bench at 2 ranks with and without the overlap
set -x
O=gpurun_out/r2c6
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout -k 10 600 python -m pytest tests/test_dist_gpu.py -q -m gpu --tb=short -rA -p no:cacheprovider > $O/dist_test.log 2>&1
tail -15 $O/dist_test.log
T0=$(date +%s)
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "rc=$? wall=$(( $(date +%s) - T0 ))s" | tee $O/bench_2gpu.rc; tail -c 700 $O/bench_2gpu.json; tail -5 $O/bench_2gpu.err
T0=$(date +%s)
TVTS_OVERLAP_ALLREDUCE=0 timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu_nooverlap.json 2> $O/bench_2gpu_nooverlap.err
echo "rc=$? wall=$(( $(date +%s) - T0 ))s" | tee $O/bench_2gpu_nooverlap.rc; tail -c 500 $O/bench_2gpu_nooverlap.json; tail -3 $O/bench_2gpu_nooverlap.err
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.json
