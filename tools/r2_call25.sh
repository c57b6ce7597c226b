#!/bin/bash
# round 2, GPU call 25 (1 GPU): verification of the final code the way the driver runs it + ncu --set full of the final attention kernels +
# launch list of one c3 step
set -x
O=gpurun_out/r2c25
mkdir -p $O
T0=$(date +%s); timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/gpu_suite.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 ))s" | tee $O/rc.txt; tail -3 $O/gpu_suite.log
T0=$(date +%s); timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -1 $O/smoke.log
T0=$(date +%s); timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt; tail -c 1200 $O/bench_default.json
T0=$(date +%s); timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$? wall=$(( $(date +%s) - T0 ))s" | tee -a $O/rc.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --gemm-breakdown > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 300 $O/bench_c3.json
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eager-baseline --no-graph"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_kernel -s 108 -c 2 -o $O/prof_attn_tc_bwd $B > $O/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd_kernel -s 120 -c 2 -o $O/prof_attn_tc_fwd $B > $O/ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file $O/c3_launches.csv $B > $O/c3_ncu.log 2>&1
python tools/launch_summary.py $O/c3_launches.csv > $O/c3_launch_summary.txt 2>&1; head -14 $O/c3_launch_summary.txt
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.json
timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 300 $O/bench_c5.json
