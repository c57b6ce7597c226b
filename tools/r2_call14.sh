#!/bin/bash
# round 2, GPU call 14 (1 GPU): tcgen05 attention with per-column CLS tickets / smem hand-off / early dV epilogue / vectorised bias sums,
# A/B against the previous build (build_ab/prev_fp16.so) on the same box
set -x
O=gpurun_out/r2c14
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" -p no:cacheprovider > $O/attn_tests.log 2>&1; echo "attn tests rc=$?" | tee $O/rc.txt; tail -3 $O/attn_tests.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainstep_gpu.py -q -m gpu -x -p no:cacheprovider > $O/model_tests.log 2>&1; echo "model tests rc=$?" | tee -a $O/rc.txt; tail -3 $O/model_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline"
for i in 1 2; do
  TVTS_LIB_PATH=build_ab/prev_fp16.so timeout 300 $B > $O/bench_prev_$i.json 2> $O/bench_prev_$i.err; tail -c 200 $O/bench_prev_$i.json
  timeout 300 $B > $O/bench_new_$i.json 2> $O/bench_new_$i.err; tail -c 200 $O/bench_new_$i.json
done
N="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-eager-baseline --no-graph"
TVTS_LIB_PATH=build_ab/prev_fp16.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_tc -c 400 --csv --log-file $O/attn_prev.csv $N > $O/ncu_prev.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_tc -c 400 --csv --log-file $O/attn_new.csv $N > $O/ncu_new.log 2>&1
python - <<'PY'
import csv, collections
for tag in ("prev", "new"):
    rows = [r for r in csv.DictReader(l for l in open(f"gpurun_out/r2c14/attn_{tag}.csv") if not l.startswith("=="))]
    agg = collections.defaultdict(list)
    for r in rows[len(rows) * 3 // 4:]:
        agg[(r["Kernel Name"][:40], r["Grid Size"])].append(float(r["Metric Value"]) / 1e3)
    for k, v in sorted(agg.items()):
        print(tag, k, len(v), "mean us %.1f" % (sum(v) / len(v)))
PY
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
