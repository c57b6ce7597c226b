#!/bin/bash
# round 2, GPU call 31 (1 GPU): 100-step loss trajectories of the FINAL code against the CPU oracle + restated AdamW (north star: 1e-3)
O=gpurun_out/r2c31
mkdir -p $O
timeout 400 python tools/loss_parity.py 100 c1 > $O/lp_c1_100.log 2>&1; tail -2 $O/lp_c1_100.log
timeout 500 python tools/loss_parity.py 100 c3s > $O/lp_c3s_100.log 2>&1; tail -2 $O/lp_c3s_100.log
