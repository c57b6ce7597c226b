#!/bin/bash
# round 2, GPU call 18 (1 GPU): persistent tcgen05 attention kernels -- parity tests, phase timing, A/B against build_ab/prev_fp16.so
set -x
O=gpurun_out/r2c18
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" -p no:cacheprovider > $O/attn_tests.log 2>&1; echo "attn tests rc=$?" | tee $O/rc.txt; tail -5 $O/attn_tests.log
for m in 1 2; do TVTS_LIB_PATH=build_ab/prof_fp16.so PYTHONPATH=. timeout 300 python tools/attn_phase_prof.py $m > $O/phase_mode$m.txt 2>&1; cat $O/phase_mode$m.txt; done
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainstep_gpu.py -q -m gpu -x -p no:cacheprovider > $O/model_tests.log 2>&1; echo "model tests rc=$?" | tee -a $O/rc.txt; tail -3 $O/model_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline"
TVTS_LIB_PATH=build_ab/prev_fp16.so timeout 300 $B > $O/bench_prev_1.json 2> $O/bench_prev_1.err; tail -c 200 $O/bench_prev_1.json
timeout 300 $B > $O/bench_new_1.json 2> $O/bench_new_1.err; tail -c 200 $O/bench_new_1.json
N="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eager-baseline --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_tc -c 400 --csv --log-file $O/attn_new.csv $N > $O/ncu_new.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.DictReader(l for l in open("gpurun_out/r2c18/attn_new.csv") if not l.startswith("=="))]
agg = collections.defaultdict(list)
for r in rows[len(rows) * 3 // 4:]:
    agg[(r["Kernel Name"][:40], r["Grid Size"])].append(float(r["Metric Value"]) / 1e3)
for k, v in sorted(agg.items()):
    print("new", k, len(v), "mean us %.1f" % (sum(v) / len(v)))
PY
