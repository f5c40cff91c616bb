"""Entropy reduction and OoD detection metrics (mirrors `runia_core.evaluation` for the hot path)."""
from . import entropy, metrics
from .entropy import *  # noqa: F401,F403
from .metrics import *  # noqa: F401,F403

__all__ = []
__all__ += entropy.__all__
__all__ += metrics.__all__
