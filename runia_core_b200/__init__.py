"""runia_core_b200 -- B200-native (sm_100a) implementation of RunIA-core's post-hoc OoD scoring hot
path behind the reference's own Python API (`get_dl_h_z`, `apply_pca_ds_split` /
`apply_pca_transform`, and the `setup()` / `postprocess()` interface of the registered
postprocessors).  Import paths mirror `runia_core`:

    from runia_core_b200.evaluation import get_dl_h_z
    from runia_core_b200 import apply_pca_ds_split, apply_pca_transform
    from runia_core_b200.inference import postprocessors_dict, MDLatentSpace, LaREMPostprocessor

`install_as_runia_core()` registers the package under the name `runia_core` for callers that
import the reference's paths."""
import sys as _sys

from . import _lib  # noqa: F401  (fails loudly when libruniab200.so is missing)
from . import dimensionality_reduction, evaluation, feature_extraction, inference, llm_uncertainty
from .dimensionality_reduction import *  # noqa: F401,F403

__version__ = "0.1.0"
__all__ = ["evaluation", "feature_extraction", "inference", "llm_uncertainty", "install_as_runia_core"]
__all__ += dimensionality_reduction.__all__


def install_as_runia_core():
    """Makes `import runia_core`, `runia_core.inference`, `runia_core.evaluation`,
    `runia_core.inference.postprocessors`, ... resolve to this package."""
    me = _sys.modules[__name__]
    _sys.modules["runia_core"] = me
    for name, mod in list(_sys.modules.items()):
        if name.startswith(__name__ + "."):
            _sys.modules["runia_core" + name[len(__name__):]] = mod
    return me
