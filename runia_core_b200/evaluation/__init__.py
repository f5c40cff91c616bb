"""Entropy reduction (mirrors `runia_core.evaluation` for the hot path)."""
from . import entropy
from .entropy import *  # noqa: F401,F403

__all__ = []
__all__ += entropy.__all__
