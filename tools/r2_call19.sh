#!/bin/bash
# round 2, GPU call 19: event-timed A/B of the attention kernels alone (previous non-persistent build vs the persistent one)
O=gpurun_out/r2c19
mkdir -p $O
echo "== previous build (one CTA per tile)" | tee $O/attn_bench.txt
TVTS_LIB_PATH=build_ab/prev_fp16.so PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
echo "== persistent build" | tee -a $O/attn_bench.txt
PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
echo "== persistent build, profiling variant (CTA residency)" | tee -a $O/attn_bench.txt
TVTS_LIB_PATH=build_ab/prof_fp16.so PYTHONPATH=. timeout 300 python tools/attn_bench.py 2>&1 | tee -a $O/attn_bench.txt
