"""Drop-in for v2/model/loss.py (NormSoftmaxLoss :5-25)."""
from tvts_b200.modules import NormSoftmaxLoss  # noqa: F401
