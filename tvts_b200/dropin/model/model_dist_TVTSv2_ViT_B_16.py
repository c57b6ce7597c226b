"""Drop-in for v2/model/model_dist_TVTSv2_ViT_B_16.py (class TVTSv2_B_16 :10-116, sim_matrix :119-127)."""
from tvts_b200.modules import TVTSv2_B_16, sim_matrix  # noqa: F401
