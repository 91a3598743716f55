"""Drop-in mirror of the ``loco_rl`` API surface that sits on the LocoTouch hot path:
``loco_rl.algorithms.PPO``, ``loco_rl.storage.RolloutStorage``, ``loco_rl.modules.ActorCritic`` and the
``loco_rl.models`` building blocks of the student.  Same names, constructor arguments, attributes and error
behaviour as the reference fork of rsl_rl 2.2.4; the arithmetic runs in the sm_100a kernels of this package."""
from .algorithms import PPO  # noqa: F401
from .modules import ActorCritic, ActorCriticPreEncoderRNNEncoder, ActorCriticRecurrent, ActorCriticRNNEncoder  # noqa: F401
from .storage import RolloutStorage  # noqa: F401
