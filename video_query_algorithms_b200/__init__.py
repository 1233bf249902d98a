"""B200-native match-scoring path of Video Query (drop-in for the reference's `models` package:
`Ticket`, `TargetClip`, `Hyperparameter`, `compute_matches`), backed by libvq_b200 (include/vq.h)."""
from ._ffi import VQError
from .compute_matches import compute_matches
from .hyperparameter import Hyperparameter, resample_labelled
from .store import FeatureStore, invalidate, loss_grid, register_store
from .target_clip import TargetClip
from .ticket import Ticket

__all__ = ["VQError", "compute_matches", "Hyperparameter", "resample_labelled", "FeatureStore",
           "invalidate", "loss_grid", "register_store", "TargetClip", "Ticket"]
