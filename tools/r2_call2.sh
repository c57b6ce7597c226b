#!/bin/bash
# round 2, GPU call 2: tcgen05 attention bring-up (A/B against the mma.sync kernels), whole suite under the fp16 default, benches, ncu
set -x
O=gpurun_out/r2c2
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_attention and not window" --tb=short -rA -p no:cacheprovider -x > $O/attn_tc_tests.log 2>&1
TC_RC=$?
tail -15 $O/attn_tc_tests.log
if [ $TC_RC -ne 0 ]; then
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "test_attention and not window" --tb=line -rA -p no:cacheprovider > $O/attn_tc_tests_all.log 2>&1
  grep -E "^(PASSED|FAILED)" $O/attn_tc_tests_all.log | awk '{print $1}' | sort | uniq -c
  grep -E "^FAILED" $O/attn_tc_tests_all.log | head -40
  export TVTS_ATTN_TC=0
fi
timeout 1500 python -m pytest tests -q -m gpu --tb=short -rA -p no:cacheprovider > $O/gpu_suite.log 2>&1
tail -12 $O/gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --gemm-breakdown > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 1500 $O/bench_c3.json; tail -3 $O/bench_c3.err
TVTS_ATTN_TC=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_c3_notc.json 2> $O/bench_c3_notc.err; tail -c 600 $O/bench_c3_notc.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --trim-text > $O/bench_c3_trim.json 2> $O/bench_c3_trim.err; tail -c 600 $O/bench_c3_trim.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/c3_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/c3_ncu.log 2>&1
python tools/launch_summary.py $O/c3_launches.csv > $O/c3_launch_summary.txt 2>&1; head -40 $O/c3_launch_summary.txt
if [ -z "$TVTS_ATTN_TC" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd_kernel -s 30 -c 1 -o $O/prof_attn_tc_fwd python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_fwd.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_kernel -s 30 -c 1 -o $O/prof_attn_tc_bwd python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/ncu_bwd.log 2>&1
fi
timeout 500 python tools/loss_parity.py 100 c3s > $O/lp_c3s_100.log 2>&1; tail -2 $O/lp_c3s_100.log
timeout 500 python tools/loss_parity.py 100 c1 > $O/lp_c1_100.log 2>&1; tail -2 $O/lp_c1_100.log
timeout 300 python tools/loss_parity.py 100 tiny > $O/lp_tiny_100.log 2>&1; tail -2 $O/lp_tiny_100.log
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-eager-baseline > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 700 $O/bench_c4.json; tail -2 $O/bench_c4.err
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 900 $O/bench_c5.json; tail -2 $O/bench_c5.err
