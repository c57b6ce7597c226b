#!/bin/bash
# round 2, GPU call 24 (1 GPU): GEMM epilogue storing straight from registers (STG.256) vs TMA-store boxes -- parity in both modes, step A/B with
# the per-shape GEMM table
set -x
O=gpurun_out/r2c24
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "gemm" -p no:cacheprovider > $O/gemm_tests.log 2>&1; echo "gemm tests rc=$?" | tee $O/rc.txt; tail -3 $O/gemm_tests.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainstep_gpu.py -q -m gpu -x -p no:cacheprovider > $O/model_tests.log 2>&1; echo "model tests rc=$?" | tee -a $O/rc.txt; tail -3 $O/model_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --gemm-breakdown"
for i in 1 2; do
  TVTS_GEMM_EPI_DIRECT=0 timeout 300 $B > $O/bench_tma_$i.json 2> $O/bench_tma_$i.err; tail -c 200 $O/bench_tma_$i.json
  timeout 300 $B > $O/bench_direct_$i.json 2> $O/bench_direct_$i.err; tail -c 200 $O/bench_direct_$i.json
done
grep -A16 'GEMM breakdown' $O/bench_tma_1.err
grep -A16 'GEMM breakdown' $O/bench_direct_1.err
