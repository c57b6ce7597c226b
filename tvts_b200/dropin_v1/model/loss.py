"""Drop-in for v1/model/loss.py (NormSoftmaxLoss)."""
from tvts_b200.modules import NormSoftmaxLoss  # noqa: F401
