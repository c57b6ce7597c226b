"""Drop-in for the classes of v2/model/video_encoder_ViT_H_14.py that the H/14 model file imports (:8-9 of
model_dist_TVTSv2_ViT_H_14.py): the video ViT (:303-484), its block (:210-254), VarAttention, LayerNorm, QuickGELU."""
from tvts_b200.modules import (LayerNorm, QuickGELU, VarAttention,  # noqa: F401
                               ResidualSpaceTimeAttentionBlockH14 as ResidualSpaceTimeAttentionBlock,
                               SpaceTimeTransformerH14 as Transformer, VisionTransformerH14 as VisionTransformer)

LayerNormFp32 = LayerNorm     # the reference's fp32-upcasting variant: LayerNorm statistics are always fp32 here
