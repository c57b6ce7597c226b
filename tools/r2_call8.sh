#!/bin/bash
# round 2, GPU call 8 (1 GPU): what the driver runs at round end, on the current code: pytest -m gpu -x, smoke(), default bench.py (timed), the
# reference arm, plus the GEMM ncu capture for roofline.traffic
set -x
O=gpurun_out/r2c8
mkdir -p $O
T0=$(date +%s); timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/gpu_suite.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 ))s" | tee $O/rc.txt; tail -4 $O/gpu_suite.log
T0=$(date +%s); timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -1 $O/smoke.log
T0=$(date +%s); timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -c 1500 $O/bench_default.json
T0=$(date +%s); timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -c 600 $O/bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --gemm-breakdown > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 400 $O/bench_c3.json
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1150 -c 8 -o $O/prof_gemm $B > $O/ncu_gemm.log 2>&1
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 300 $O/bench_c5.json
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 300 $O/bench_c4.json
