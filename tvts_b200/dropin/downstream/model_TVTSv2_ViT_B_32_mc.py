"""Drop-in for v2/downstream/model_TVTSv2_ViT_B_32_mc.py (multiple-choice variant: text embeddings stay [n_choices, B, E], :62-64)."""
from tvts_b200.modules import TVTSv2_B_32_downstream_mc as TVTSv2_B_32, sim_matrix  # noqa: F401
