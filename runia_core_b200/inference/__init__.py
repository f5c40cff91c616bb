"""Scoring operators (mirrors `runia_core.inference` for the hot path)."""
from . import abstract_classes, funcs, image_level, postprocessors
from .abstract_classes import *  # noqa: F401,F403
from .funcs import *  # noqa: F401,F403
from .image_level import *  # noqa: F401,F403
from .funcs import normalizer  # noqa: F401
from .postprocessors import *  # noqa: F401,F403

__all__ = []
__all__ += abstract_classes.__all__
__all__ += postprocessors.__all__
__all__ += funcs.__all__
__all__ += image_level.__all__
