"""Drop-in for v1/model/video_encoder.py (PatchEmbed :78-99, Attention :30-56, Block :59-75, VisionTransformer :102-226)."""
from tvts_b200.modules_v1 import Attention, Block, Mlp, PatchEmbed, VisionTransformer  # noqa: F401
