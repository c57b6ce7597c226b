"""Loss-trajectory parity on the GPU (north star: step losses track the reference over many optimizer steps): 12 steps of the
toy model and 6 steps of BASELINE.json configs[0] (ViT-B/32, 2 frames, 4 pairs), both at the reference learning rates, through
TrainStep (CUDA graph replay + fused AdamW) against the CPU oracle + restated transformers.AdamW on the same batches.
tools/loss_parity.py is the long (100-step) version.  Tolerance: bf16 GEMM operands -> 5e-2 on each loss at the toy widths (the
single-step model tests hold the same bound; the GPU smoke check observed 1.4e-3 / 3.2e-3).  (These two tests were added after
the round's GPU budget was spent, hence the conservative bound; tools/loss_parity.py prints the actual deviations.)"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

from tvts_b200._lib import OPERAND  # noqa: E402

# bf16 operands: the conservative bound described above; fp16 operands (TVTS_OPERAND=fp16): the build meant to meet the north star's
# 1e-3 (emulation: 5e-4 over 20 toy steps, 3e-4 on c1) -- held to 2e-3 until its first GPU run
BOUND = 5e-2 if OPERAND == "bf16" else 2e-3


@pytest.mark.timeout(600)
def test_toy_model_12_steps():
    import loss_parity
    d1, d2, dt = loss_parity.run(12, "tiny", verbose=False)
    assert d1 < BOUND and d2 < BOUND, (d1, d2)


@pytest.mark.timeout(900)
def test_c1_6_steps():
    import loss_parity
    d1, d2, dt = loss_parity.run(6, "c1", verbose=False)
    assert d1 < BOUND and d2 < BOUND, (d1, d2)
