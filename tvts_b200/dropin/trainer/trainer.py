"""Drop-in for the hot-path pieces of v2/trainer/trainer.py: AllGather_multi (:41-57), the step body of
Trainer_TVTSv2_*._train_epoch (:463-499) as `TrainStep`, and the epoch loop / validation (`Trainer_TVTSv2_*`, :361-635)."""
from tvts_b200.trainer import (AllGather_multi, TrainStep, Trainer_TVTSv2, Trainer_TVTSv2_B_16, Trainer_TVTSv2_B_32, Trainer_TVTSv2_H_14,  # noqa: F401
                               gather_embeddings, validate, verbose, format_nested_metrics_for_writer)
