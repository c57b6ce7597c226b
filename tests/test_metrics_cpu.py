"""tvts_b200.metrics against the unmodified reference metric functions (v2/model/metric.py) when /root/reference is available,
and against hand-checked small cases otherwise."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from tvts_b200 import metrics as MT

REF = "/root/reference/v2/model/metric.py"


def _load_reference():
    for name in ("ipdb",):                       # debug-only import of the reference module, absent from this image
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_metric", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_small_hand_checked_case():
    sims = np.array([[0.9, 0.1, 0.0], [0.2, 0.1, 0.3], [0.0, 0.5, 0.6]])
    m = MT.t2v_metrics(sims)
    assert m["R1"] == pytest.approx(100 * 2 / 3) and m["MedR"] == 1.0 and m["MeanR"] == pytest.approx((1 + 3 + 1) / 3)
    m = MT.v2t_metrics(sims)
    assert m["R1"] == pytest.approx(100 * 2 / 3)


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference checkout not present")
def test_matches_reference_metric_functions():
    ref = _load_reference()
    rs = np.random.RandomState(0)
    for nv, qpv in ((32, 1), (20, 3), (64, 1)):
        sims = rs.randn(nv * qpv, nv).astype(np.float32)
        sims[np.arange(nv * qpv), np.arange(nv * qpv) // qpv] += 1.0
        a, b = MT.t2v_metrics(sims), ref.t2v_metrics(sims)
        for k in ("R1", "R5", "R10", "R50", "MedR", "MeanR"):
            assert a[k] == pytest.approx(float(b[k])), (k, a[k], b[k])
        a, b = MT.v2t_metrics(sims), ref.v2t_metrics(sims)
        for k in ("R1", "R5", "R10", "R50", "MedR", "MeanR"):
            assert a[k] == pytest.approx(float(b[k])), (k, a[k], b[k])
