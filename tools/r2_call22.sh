#!/bin/bash
# round 2, GPU call 22: phase timing with per-CTA spans (globaltimer) of the persistent attention kernels
O=gpurun_out/r2c22
mkdir -p $O
for m in 1 2; do TVTS_LIB_PATH=build_ab/prof_fp16.so PYTHONPATH=. timeout 300 python tools/attn_phase_prof.py $m > $O/phase_mode$m.txt 2>&1; cat $O/phase_mode$m.txt; done
