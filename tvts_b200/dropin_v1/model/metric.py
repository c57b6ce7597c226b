"""Drop-in for model/metric.py: the retrieval metrics every shipped config names (t2v_metrics :16-125, v2t_metrics :127-217,
cols2metrics :285-295); same numbers as the reference's functions (tests/test_metrics_cpu.py)."""
from tvts_b200.metrics import cols2metrics, t2v_metrics, v2t_metrics  # noqa: F401
