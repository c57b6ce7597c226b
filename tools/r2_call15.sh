#!/bin/bash
# round 2, GPU call 15: phase timing of the tcgen05 attention kernels (library variant built with -DTVTS_ATTN_PROF:
#   nvcc ... -DTVTS_OPERAND_FP16 --use_fast_math -DTVTS_ATTN_PROF -c tvts_b200/csrc/attention_tc.cu -o build_ab/obj/attention_tc_prof.o;
#   nvcc -shared -o build_ab/prof_fp16.so <the other build_fp16 objects> build_ab/obj/attention_tc_prof.o)
O=gpurun_out/r2c15
mkdir -p $O
for m in 1 2; do TVTS_LIB_PATH=build_ab/prof_fp16.so PYTHONPATH=. timeout 300 python tools/attn_phase_prof.py $m > $O/phase_mode$m.txt 2>&1; cat $O/phase_mode$m.txt; done
