#!/bin/bash
# round 2, GPU call 9 (2 GPUs): the fused InfoNCE + sort-CE launch -- kernel tests, whole suite incl. the 2-rank NCCL test, bench with / without it
set -x
O=gpurun_out/r2c9
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fused" --tb=short -rA -p no:cacheprovider > $O/fused_tests.log 2>&1; tail -12 $O/fused_tests.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $O/gpu_suite.log 2>&1; tail -5 $O/gpu_suite.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.json; tail -3 $O/bench_c3.err
CUDA_VISIBLE_DEVICES=0 TVTS_FUSED_LOSS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_unfused.json 2> $O/bench_c3_unfused.err; tail -c 300 $O/bench_c3_unfused.json
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?"; tail -c 300 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
CUDA_VISIBLE_DEVICES=0 timeout 400 python tools/loss_parity.py 100 c1 > $O/lp_c1_100.log 2>&1; tail -2 $O/lp_c1_100.log
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/c3_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/c3_ncu.log 2>&1
python tools/launch_summary.py $O/c3_launches.csv > $O/c3_launch_summary.txt 2>&1; grep -n "fused\|launches in capture" $O/c3_launch_summary.txt
