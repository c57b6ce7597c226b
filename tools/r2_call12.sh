#!/bin/bash
# round 2, GPU call 12 (1 GPU): wave-aware solo / pair GEMM tile choice -- A/B with the per-shape breakdown
set -x
O=gpurun_out/r2c12
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > $O/tests.log 2>&1; tail -3 $O/tests.log
for P in 0 115 107; do
  TVTS_GEMM_SOLO_PENALTY=$P timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --gemm-breakdown > $O/bench_c3_p$P.json 2> $O/bench_c3_p$P.err
  tail -c 250 $O/bench_c3_p$P.json; grep -A14 "GEMM breakdown" $O/bench_c3_p$P.err
done
TVTS_GEMM_SOLO_PENALTY=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_p0_b.json 2> /dev/null; tail -c 250 $O/bench_c3_p0_b.json
TVTS_GEMM_SOLO_PENALTY=115 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_p115_b.json 2> /dev/null; tail -c 250 $O/bench_c3_p115_b.json
