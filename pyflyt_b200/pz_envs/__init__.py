"""Vectorised mirrors of ``PyFlyt.pz_envs`` (hot-path rows of SURVEY.md §8 only)."""
from .ma_fixedwing_dogfight_env import MAFixedwingDogfightVecEnv  # noqa: F401
from .ma_fixedwing_dogfight_split import MAFixedwingDogfightSplitEnv, split_agent_range, spawn_poses  # noqa: F401
from .ma_quadx_hover_env import MAQuadXHoverVecEnv  # noqa: F401
