#!/bin/bash
# First GPU call of round 2: everything that was written after the round-1 GPU budget ran out, in one box session.
#   /usr/local/graft/bin/gpurun --timeout 3300 -- 'bash tools/round2_bringup.sh'          (~40-50 min of box time; every step has
#   its own `timeout`; run sections 1-2 alone first -- `bash tools/round2_bringup.sh quick` -- if the budget is tight)
# Results land in gpurun_out/r2_bringup/.
set -x
mkdir -p gpurun_out/r2_bringup
O=gpurun_out/r2_bringup
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
# 1. staged tests (H/14 attention d=80, padded patch embed, v1 kernels + models, downstream): run WITHOUT the xfail veil
TVTS_RUN_STAGED=1 timeout 900 python -m pytest tests/test_zz_round1_unverified_gpu.py -q -m gpu -x --tb=short > $O/zz_first_failure.log 2>&1
TVTS_RUN_STAGED=1 timeout 900 python -m pytest tests/test_zz_round1_unverified_gpu.py -q -m gpu --tb=line -rA > $O/zz_all.log 2>&1
# 2. the verified suite (must stay green) incl. the tests that were never run on a GPU in round 1
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_zy_staged_wrapper_gpu.py > $O/gpu_suite.log 2>&1
if [ "$1" = "quick" ]; then tail -5 $O/zz_all.log $O/gpu_suite.log; exit 0; fi
# 3. loss trajectories (north star: 100 steps within 1e-3)
timeout 900 python tools/loss_parity.py 100 c1 > $O/loss_parity_c1_100.log 2>&1
timeout 300 python tools/loss_parity.py 100 tiny > $O/loss_parity_tiny_100.log 2>&1
timeout 600 python tools/loss_parity.py 30 tiny_h > $O/loss_parity_tiny_h_30.log 2>&1
# 3b. the fp16-operand build (libtvts_b200_fp16.so): kernel / model / train-step suites, then the same trajectories and the headline bench
TVTS_OPERAND=fp16 timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_trainstep_gpu.py tests/test_y_late_additions_gpu.py -q -m gpu --tb=line -rA > $O/fp16_suites.log 2>&1
TVTS_OPERAND=fp16 timeout 900 python tools/loss_parity.py 100 c1 > $O/fp16_loss_parity_c1_100.log 2>&1
TVTS_OPERAND=fp16 timeout 300 python tools/loss_parity.py 100 tiny > $O/fp16_loss_parity_tiny_100.log 2>&1
TVTS_OPERAND=fp16 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/fp16_bench_c3.json 2> $O/fp16_bench_c3.err
# 4. benches: headline, then H/14
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 --u8-input --no-cpu-baseline > $O/bench_c3_u8.json 2> $O/bench_c3_u8.err
timeout 600 python bench.py --steps 20 --warmup 5 --trim-text --no-cpu-baseline > $O/bench_c3_trim.json 2> $O/bench_c3_trim.err
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
tail -3 $O/*.log $O/*.json
timeout 600 python tools/bench_v1.py --steps 5 --warmup 3 > $O/bench_v1_c5.json 2> $O/bench_v1_c5.err
tail -2 $O/bench_v1_c5.json $O/bench_v1_c5.err
# 5. launch lists (cold-cache, serialised: shares only) of the H/14 step and of the trimmed-text c3 step
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/c4_launches.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/c4_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/c3_trim_launches.csv python bench.py --trim-text --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/c3_trim_ncu.log 2>&1
python tools/launch_summary.py $O/c4_launches.csv > $O/c4_launch_summary.txt 2>&1
python tools/launch_summary.py $O/c3_trim_launches.csv > $O/c3_trim_launch_summary.txt 2>&1
