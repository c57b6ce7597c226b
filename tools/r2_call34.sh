#!/bin/bash
# round 2, GPU call 34 (2 GPUs): tvts_comm_* (the C ABI's own NCCL communicator, TVTS_COMM=native) inside the 2-rank parity test
O=gpurun_out/r2c34
mkdir -p $O
timeout -k 10 400 python -m pytest tests/test_dist_gpu.py -q -m gpu --tb=short -rA -p no:cacheprovider > $O/dist_test.log 2>&1; echo "dist test rc=$?" | tee $O/rc.txt
tail -25 $O/dist_test.log
