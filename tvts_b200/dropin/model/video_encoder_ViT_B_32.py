"""Drop-in for v2/model/video_encoder_ViT_B_32.py (:18-235)."""
from tvts_b200.modules import (LayerNorm, QuickGELU, ResidualSpaceTimeAttentionBlock, SpaceTimeTransformer, VarAttention,  # noqa: F401
                               VisionTransformer)
