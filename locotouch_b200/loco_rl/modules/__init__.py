from .actor_critic import ActorCritic  # noqa: F401
from .actor_critic_recurrent import ActorCriticPreEncoderRNNEncoder, ActorCriticRecurrent, ActorCriticRNNEncoder, Memory  # noqa: F401
