"""Reducers on the hot path that feed the entropy estimator: `get_mean_or_fullmean_ls_sample`
(`runia_core.feature_extraction.utils`), `MCSamplerModule` (`...abstract_classes`) and the RoI reducers of the
object-level path (`...object_level`); hooks, extractors and model wrappers are out of scope."""
from . import abstract_classes, object_level, utils
from .abstract_classes import *  # noqa: F401,F403
from .object_level import _dropblock_rois_get_entropy, _reduce_features_to_rois  # noqa: F401
from .utils import *  # noqa: F401,F403

__all__ = []
__all__ += utils.__all__
__all__ += abstract_classes.__all__
