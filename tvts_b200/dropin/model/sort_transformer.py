"""Drop-in for v2/model/sort_transformer.py (:16-142)."""
from tvts_b200.modules import AttnBlock, Mlp, SelfAttention, SortTransformer  # noqa: F401
