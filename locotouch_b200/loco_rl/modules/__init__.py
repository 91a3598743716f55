from .actor_critic import ActorCritic  # noqa: F401
