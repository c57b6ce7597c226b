"""Drop-in for the hot-path pieces of v2/trainer/trainer.py: AllGather_multi (:41-57) and the body of
Trainer_TVTSv2_*._train_epoch (:463-499) as `TrainStep` (tokenised batch in, losses out)."""
from tvts_b200.trainer import AllGather_multi, TrainStep, gather_embeddings  # noqa: F401
