"""LLM uncertainty scores on the hot path (mirrors `runia_core.llm_uncertainty` for `eigen_score`)."""
from . import scores
from .scores import *  # noqa: F401,F403

__all__ = []
__all__ += scores.__all__
