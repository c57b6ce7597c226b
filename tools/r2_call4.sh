#!/bin/bash
# round 2, GPU call 4: LN backward single-phase rows, L2 prefetch in the tcgen05 attention, text tower on a second stream, trimmed text by default
set -x
O=gpurun_out/r2c4
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --tb=short -rA -p no:cacheprovider -x > $O/gpu_suite.log 2>&1
tail -6 $O/gpu_suite.log
timeout 700 python bench.py --steps 20 --warmup 5 --gemm-breakdown > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 1200 $O/bench_c3.json; tail -3 $O/bench_c3.err
TVTS_TEXT_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_onestream.json 2> $O/bench_c3_onestream.err; tail -c 300 $O/bench_c3_onestream.json
TVTS_ATTN_TC_PREFETCH=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_noprefetch.json 2> $O/bench_c3_noprefetch.err; tail -c 300 $O/bench_c3_noprefetch.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-trim-text > $O/bench_c3_notrim.json 2> $O/bench_c3_notrim.err; tail -c 300 $O/bench_c3_notrim.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/c3_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/c3_ncu.log 2>&1
python tools/launch_summary.py $O/c3_launches.csv > $O/c3_launch_summary.txt 2>&1; head -24 $O/c3_launch_summary.txt
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_kernel -s 240 -c 1 -o $O/prof_ln_bwd $B > $O/ncu3.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_kernel -s 108 -c 2 -o $O/prof_attn_tc_bwd $B > $O/ncu1.log 2>&1
timeout 500 python tools/loss_parity.py 100 c3s > $O/lp_c3s_100.log 2>&1; tail -2 $O/lp_c3s_100.log
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-eager-baseline --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 400 $O/bench_c4.json; tail -2 $O/bench_c4.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 400 $O/bench_c2.json; tail -2 $O/bench_c2.err
