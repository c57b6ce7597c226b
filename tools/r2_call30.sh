#!/bin/bash
# round 2, GPU call 30 (1 GPU): last verification of the committed code the way the driver runs it, and the ncu --set full capture of the
# GEMM kernel as shipped that roofline.traffic in bench.py quotes
set -x
O=gpurun_out/r2c30
mkdir -p $O
T0=$(date +%s); timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/gpu_suite.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 ))s" | tee $O/rc.txt; tail -3 $O/gpu_suite.log
T0=$(date +%s); timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -1 $O/smoke.log
T0=$(date +%s); timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -c 600 $O/bench_default.json
T0=$(date +%s); timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eager-baseline --no-graph"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1150 -c 8 -o $O/prof_gemm $B > $O/ncu_gemm.log 2>&1
ls -la $O
