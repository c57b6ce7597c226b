"""Drop-in for v1/model/sort_transformer.py (same SortTransformer as v2; v1 uses embed_dim 768, 12 heads)."""
from tvts_b200.modules import AttnBlock, Mlp, SelfAttention, SortTransformer  # noqa: F401
