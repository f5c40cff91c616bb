"""Reducers on the hot path that feed the entropy estimator (mirrors `runia_core.feature_extraction.utils`
for `get_mean_or_fullmean_ls_sample`; hooks, samplers and model wrappers are out of scope)."""
from . import utils
from .utils import *  # noqa: F401,F403

__all__ = []
__all__ += utils.__all__
