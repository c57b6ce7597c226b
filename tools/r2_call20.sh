#!/bin/bash
# round 2, GPU call 20 (1 GPU): persistent tcgen05 attention with the dynamic tile scheduler -- parity, event-timed A/B, step A/B
set -x
O=gpurun_out/r2c20
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" -p no:cacheprovider > $O/attn_tests.log 2>&1; echo "attn tests rc=$?" | tee $O/rc.txt; tail -3 $O/attn_tests.log
echo "== previous build (one CTA per tile)" | tee $O/attn_bench.txt
TVTS_LIB_PATH=build_ab/prev_fp16.so PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
echo "== persistent build" | tee -a $O/attn_bench.txt
PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainstep_gpu.py -q -m gpu -x -p no:cacheprovider > $O/model_tests.log 2>&1; echo "model tests rc=$?" | tee -a $O/rc.txt; tail -3 $O/model_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline"
TVTS_LIB_PATH=build_ab/prev_fp16.so timeout 300 $B > $O/bench_prev_1.json 2> $O/bench_prev_1.err; tail -c 200 $O/bench_prev_1.json
timeout 300 $B > $O/bench_new_1.json 2> $O/bench_new_1.err; tail -c 200 $O/bench_new_1.json
