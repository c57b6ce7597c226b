"""Drop-in for v2/model/model_dist_TVTSv2_ViT_B_32.py (class TVTSv2_B_32, sim_matrix)."""
from tvts_b200.modules import TVTSv2_B_32, sim_matrix  # noqa: F401
