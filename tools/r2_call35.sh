#!/bin/bash
# round 2, GPU call 35 (1 GPU): the whole GPU suite on the last commit (after the AdamW entry-point split and the communicator entry points)
O=gpurun_out/r2c35
mkdir -p $O
T0=$(date +%s); timeout 100 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/gpu_suite.log 2>&1; echo "pytest rc=$? wall=$(( $(date +%s) - T0 ))s" | tee $O/rc.txt; tail -3 $O/gpu_suite.log
