#!/bin/bash
# round 2, GPU call 10 (1 GPU): CLS merge folded into the tile kernels (atomic ticket), fused loss kernel v2, small_linear_bwd over rows
set -x
O=gpurun_out/r2c10
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $O/gpu_suite.log 2>&1; tail -5 $O/gpu_suite.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.json; tail -3 $O/bench_c3.err
TVTS_FUSED_LOSS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_unfused.json 2> $O/bench_c3_unfused.err; tail -c 300 $O/bench_c3_unfused.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_b.json 2> $O/bench_c3_b.err; tail -c 300 $O/bench_c3_b.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/c3_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/c3_ncu.log 2>&1
python tools/launch_summary.py $O/c3_launches.csv > $O/c3_launch_summary.txt 2>&1; head -30 $O/c3_launch_summary.txt
timeout 400 python tools/loss_parity.py 100 c1 > $O/lp_c1_100.log 2>&1; tail -2 $O/lp_c1_100.log
