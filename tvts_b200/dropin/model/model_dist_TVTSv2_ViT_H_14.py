"""Drop-in for v2/model/model_dist_TVTSv2_ViT_H_14.py (class TVTSv2_H_14 :13-158, sim_matrix :161-169)."""
from tvts_b200.modules import TVTSv2_H_14, sim_matrix  # noqa: F401
