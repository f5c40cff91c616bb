"""Reducers on the hot path that feed the entropy estimator: `get_mean_or_fullmean_ls_sample`
(`runia_core.feature_extraction.utils`) and `MCSamplerModule` (`...abstract_classes`); hooks, extractors and
model wrappers are out of scope."""
from . import abstract_classes, utils
from .abstract_classes import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403

__all__ = []
__all__ += utils.__all__
__all__ += abstract_classes.__all__
