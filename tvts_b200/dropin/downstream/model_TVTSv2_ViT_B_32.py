"""Drop-in for v2/downstream/model_TVTSv2_ViT_B_32.py (class TVTSv2_B_32 :10-98: mask_ratio 0, no sort head; sim_matrix :101-109)."""
from tvts_b200.modules import TVTSv2_B_32_downstream as TVTSv2_B_32, sim_matrix  # noqa: F401
