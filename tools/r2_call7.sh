#!/bin/bash
# round 2, GPU call 7 (8 GPUs): 2-rank NCCL parity test, then the c3 bench at 8 ranks: bucketed overlap on / off / on with fewer NCCL CTAs; 1-GPU line
set -x
O=gpurun_out/r2c7
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
CUDA_VISIBLE_DEVICES=0,1 timeout -k 10 600 python -m pytest tests/test_dist_gpu.py -q -m gpu --tb=short -rA -p no:cacheprovider > $O/dist_test.log 2>&1
tail -6 $O/dist_test.log
run8() {  # name, extra env...
  local name=$1; shift
  local T0=$(date +%s)
  env "$@" timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > $O/bench_8gpu_$name.json 2> $O/bench_8gpu_$name.err
  echo "$name rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt
  tail -c 350 $O/bench_8gpu_$name.json | head -c 350; echo
}
run8 overlap TVTS_OVERLAP_ALLREDUCE=1
run8 nooverlap TVTS_OVERLAP_ALLREDUCE=0
run8 overlap_cta8 TVTS_OVERLAP_ALLREDUCE=1 NCCL_MAX_CTAS=8
run8 overlap_cta16 TVTS_OVERLAP_ALLREDUCE=1 NCCL_MAX_CTAS=16
T0=$(date +%s)
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_8gpu_full.json 2> $O/bench_8gpu_full.err
echo "full rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt
CUDA_VISIBLE_DEVICES=0 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.json
for n in 2 4; do
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29300 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
  echo "n=$n rc=$?" | tee -a $O/rc.txt; tail -c 300 $O/bench_${n}gpu.json | head -c 300; echo
done
