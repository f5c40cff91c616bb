"""Entropy reduction, OoD detection metrics and the baselines driver (mirrors `runia_core.evaluation` for the hot path)."""
from . import baselines, entropy, metrics
from .baselines import *  # noqa: F401,F403
from .entropy import *  # noqa: F401,F403
from .metrics import *  # noqa: F401,F403

__all__ = []
__all__ += entropy.__all__
__all__ += metrics.__all__
__all__ += baselines.__all__
