from .TaskEnvironment import SpatialGoalEnvironment  # noqa: F401
