#!/bin/bash
# round 2, GPU call 11 (1 GPU): GEMM epilogue reorderings (deferred box acquire / operand re-request), fused-loss unroll
set -x
O=gpurun_out/r2c11
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" --tb=short -p no:cacheprovider -x > $O/gemm_tests.log 2>&1; tail -4 $O/gemm_tests.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $O/gpu_suite.log 2>&1; tail -4 $O/gpu_suite.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gemm-breakdown > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.json; grep -A22 "GEMM breakdown" $O/bench_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_b.json 2> $O/bench_c3_b.err; tail -c 300 $O/bench_c3_b.json
timeout 400 python tools/loss_parity.py 100 c3s > $O/lp_c3s_100.log 2>&1; tail -2 $O/lp_c3s_100.log
