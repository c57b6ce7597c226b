#!/bin/bash
# round 2, GPU call 21 (1 GPU): persistent tcgen05 attention with direct register stores -- parity, event-timed A/B (forward 4 vs 3 CTAs/SM), phases, step A/B
set -x
O=gpurun_out/r2c21
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" -p no:cacheprovider > $O/attn_tests.log 2>&1; echo "attn tests rc=$?" | tee $O/rc.txt; tail -3 $O/attn_tests.log
echo "== previous build (one CTA per tile)" | tee $O/attn_bench.txt
TVTS_LIB_PATH=build_ab/prev_fp16.so PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
echo "== persistent build, forward 4 CTAs / SM" | tee -a $O/attn_bench.txt
PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
echo "== persistent build, forward 3 CTAs / SM" | tee -a $O/attn_bench.txt
TVTS_LIB_PATH=build_ab/fwd3_fp16.so PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
for m in 1; do TVTS_LIB_PATH=build_ab/prof_fp16.so PYTHONPATH=. timeout 300 python tools/attn_phase_prof.py $m > $O/phase_mode$m.txt 2>&1; cat $O/phase_mode$m.txt; done
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_trainstep_gpu.py -q -m gpu -x -p no:cacheprovider > $O/model_tests.log 2>&1; echo "model tests rc=$?" | tee -a $O/rc.txt; tail -3 $O/model_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline"
TVTS_LIB_PATH=build_ab/prev_fp16.so timeout 300 $B > $O/bench_prev_1.json 2> $O/bench_prev_1.err; tail -c 200 $O/bench_prev_1.json
timeout 300 $B > $O/bench_new_1.json 2> $O/bench_new_1.err; tail -c 200 $O/bench_new_1.json
