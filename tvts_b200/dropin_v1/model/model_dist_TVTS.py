"""Drop-in for v1/model/model_dist_TVTS.py (class TVTS :18-141, sim_matrix :144-152)."""
from tvts_b200.modules_v1 import TVTS, sim_matrix  # noqa: F401
