"""Drop-in pieces of ``locotouch.distill`` on the hot path."""
from .batch import masked_mse_loss, pad_trajectories  # noqa: F401
from .tactile_recorder import TactileRecorder  # noqa: F401
from .cfg import DistillationCfg, DistillationRandCylinderCNNRNNMonCfg  # noqa: F401
from .student import Student  # noqa: F401
from .replay_buffer import ReplayBuffer  # noqa: F401
