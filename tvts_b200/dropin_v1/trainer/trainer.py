"""Drop-in for the hot-path pieces of v1/trainer/trainer.py: AllGather_multi (:18-37) and Trainer_TVTS (:40-260)."""
from tvts_b200.trainer import AllGather_multi, TrainStep, Trainer_TVTS, gather_embeddings, validate, verbose, format_nested_metrics_for_writer  # noqa: F401
