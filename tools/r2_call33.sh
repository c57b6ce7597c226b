#!/bin/bash
# round 2, GPU call 33 (1 GPU): the bfloat16 build (TVTS_OPERAND=bf16) of the final code -- kernel and model parity tests, one bench line
O=gpurun_out/r2c33
mkdir -p $O
TVTS_OPERAND=bf16 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x -p no:cacheprovider > $O/bf16_tests.log 2>&1; echo "bf16 tests rc=$?" | tee $O/rc.txt; tail -3 $O/bf16_tests.log
TVTS_OPERAND=bf16 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > $O/bench_c3_bf16.json 2> $O/bench_c3_bf16.err; tail -c 250 $O/bench_c3_bf16.json
