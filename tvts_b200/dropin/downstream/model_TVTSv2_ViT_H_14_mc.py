"""Drop-in for v2/downstream/model_TVTSv2_ViT_H_14_mc.py (multiple-choice variant: text embeddings stay [n_choices, B, E])."""
from tvts_b200.modules import TVTSv2_H_14_downstream_mc as TVTSv2_H_14, sim_matrix  # noqa: F401
