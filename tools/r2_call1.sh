#!/bin/bash
# round 2, GPU call 1: graph-vs-eager diagnosis, the whole GPU suite without -x (staged cases un-veiled, separate process), loss trajectories, benches
set -x
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi -L > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 400 python tools/diag_graph.py > $O/diag.log 2>&1
tail -40 $O/diag.log
timeout 1200 python -m pytest tests -q -m gpu --tb=short -rA -p no:cacheprovider --deselect tests/test_zy_staged_wrapper_gpu.py > $O/gpu_suite.log 2>&1
tail -15 $O/gpu_suite.log
TVTS_RUN_STAGED=1 timeout 900 python -m pytest tests/test_zz_round1_unverified_gpu.py -q -m gpu --tb=short -rA -p no:cacheprovider > $O/zz_all.log 2>&1
tail -30 $O/zz_all.log
timeout 300 python tools/loss_parity.py 100 tiny > $O/lp_tiny_100.log 2>&1; tail -3 $O/lp_tiny_100.log
TVTS_OPERAND=fp16 timeout 300 python tools/loss_parity.py 100 tiny > $O/lp_fp16_tiny_100.log 2>&1; tail -3 $O/lp_fp16_tiny_100.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; tail -2 $O/bench_c3.json $O/bench_c3.err
timeout 700 python tools/loss_parity.py 100 c1 > $O/lp_c1_100.log 2>&1; tail -3 $O/lp_c1_100.log
TVTS_OPERAND=fp16 timeout 700 python tools/loss_parity.py 100 c1 > $O/lp_fp16_c1_100.log 2>&1; tail -3 $O/lp_fp16_c1_100.log
timeout 400 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; tail -2 $O/bench_c4.json $O/bench_c4.err
timeout 300 python tools/bench_v1.py --steps 5 --warmup 3 > $O/bench_v1_c5.json 2> $O/bench_v1_c5.err; tail -2 $O/bench_v1_c5.json $O/bench_v1_c5.err
DIAG_PART=1 timeout 500 compute-sanitizer --tool initcheck --print-limit 30 python tools/diag_graph.py TINY_B_MASK 2 > $O/initcheck.log 2>&1; grep -c "Uninitialized" $O/initcheck.log; tail -5 $O/initcheck.log
