"""Drop-in for v2/downstream/model_TVTSv2_ViT_H_14.py (class TVTSv2_H_14: mask_ratio 0, no sort head; sim_matrix)."""
from tvts_b200.modules import TVTSv2_H_14_downstream as TVTSv2_H_14, sim_matrix  # noqa: F401
