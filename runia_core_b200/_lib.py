"""ctypes binding of libruniab200.so (C ABI declared in include/runia_b200.h).

The library is the product's only compute path: if it cannot be loaded this module raises --
there is no NumPy / PyTorch fallback behind it."""
import ctypes
import os
import shutil
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libruniab200.so")

RUNIA_OK = 0
ROWNORM_MD = 0
ROWNORM_VIM = 1


class RuniaB200Error(RuntimeError):
    pass


def _load():
    from .build import _stale, build_library

    have_nvcc = bool(shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))
    if not os.path.exists(LIB_PATH) or (_stale() and have_nvcc):
        # developer convenience: (re)build in-tree when a CUDA toolkit is present; never fall back
        if not have_nvcc:
            raise ImportError(
                f"{LIB_PATH} is missing and nvcc is not available: build it with "
                "`python -m runia_core_b200.build` (this package has no CPU fallback)")
        build_library()
    return ctypes.CDLL(LIB_PATH)


_lib = _load()

_P = c_void_p
_SIGS = {
    "runia_b200_abi_version": (c_int, []),
    "runia_b200_last_error": (c_char_p, []),
    "runia_b200_launch_count": (c_int64, []),
    "runia_mcd_entropy_f32": (c_int, [_P, c_int64, c_int, c_int, c_int, c_double, c_double, _P, _P, _P]),
    "runia_center_cast": (c_int, [_P, c_int, c_int64, c_int, _P, _P, _P]),
    "runia_pca_transform_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P, _P]),
    "runia_rownorm_score_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, c_int, _P, c_int, c_float,
                                        _P, _P, _P]),
    "runia_classcond_mahalanobis_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P, _P, c_int, _P, _P, _P]),
    "runia_gmm_lse_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, c_int, _P, _P]),
    "runia_classcond_mahalanobis_tc": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, _P, _P, _P, c_int, _P, _P, _P]),
    "runia_gmm_lse_tc": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, _P, c_int, _P, _P]),
    "runia_normalize_rows": (c_int, [_P, c_int, c_int64, c_int, _P, _P]),
    "runia_row_sqnorm_f32": (c_int, [_P, c_int64, c_int, _P, _P]),
    "runia_knn_workspace_bytes": (c_int64, [c_int64, c_int64, c_int, c_int]),
    "runia_knn_search_f32": (c_int, [_P, c_int64, _P, _P, _P, _P, c_int64, c_int, c_int, c_int64, _P, _P, _P, _P,
                                     _P, _P, c_int64, _P]),
    "runia_knn_search_ex_f32": (c_int, [_P, c_int64, _P, _P, _P, _P, c_int64, c_int, c_int, c_int64, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "runia_split_tf32": (c_int, [_P, c_int64, _P, _P, _P]),
    "runia_rownorm_score_tc": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, _P, c_int, _P, c_int, c_float,
                                       _P, _P, _P]),
    "runia_pca_transform_tc": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, _P, _P, _P]),
    "runia_topk_merge": (c_int, [_P, _P, c_int, c_int64, c_int, _P, _P, _P, _P]),
    "runia_kde_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "runia_kde_lse_f32": (c_int, [_P, c_int64, _P, _P, _P, c_int64, c_int, c_double, c_int64, _P, _P, _P, _P,
                                  c_int64, _P]),
    "runia_ood_metrics_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "runia_ood_metrics": (c_int, [_P, c_int64, _P, c_int64, c_int, _P, _P, _P, _P, c_int64, _P]),
    "runia_eigen_score_f32": (c_int, [_P, c_int, c_int, c_double, _P, _P]),
    "runia_pred_uncertainty_f32": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P]),
    "runia_spatial_mean_f32": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P]),
    "runia_sort_f32_workspace_bytes": (c_int64, [c_int64]),
    "runia_sort_f32": (c_int, [_P, c_int64, _P, _P, c_int64, _P]),
    "runia_class_mean_f32": (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P, _P]),
    "runia_centered_gram_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "runia_centered_gram_f64": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "runia_shifted_gram_f64": (c_int, [_P, _P, c_int64, c_int, _P, _P, _P, c_size_t, _P]),
    "runia_mc_dropblock_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "runia_mc_dropblock_mean_f32": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "runia_mc_dropblock_apply_f32": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "runia_logit_scores_f32": (c_int, [_P, c_int64, c_int, c_float, c_int, _P, _P, _P, _P]),
    "runia_clip_linear_lse_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, c_float, _P, _P]),
    "runia_tf32_peak_probe": (c_int, [c_int, _P, _P]),
    "runia_clip_linear_lse_tc": (c_int, [_P, c_int64, c_int, _P, _P, _P, c_int, c_float, _P, _P]),
    "runia_ash_linear_lse_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, c_int, _P, _P]),
    "runia_ash_prune_f32": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "runia_gen_entropy_f32": (c_int, [_P, c_int64, c_int, c_float, c_int, _P, _P]),
    "runia_linear_f32": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P]),
    "runia_stage_h2d": (c_int, [_P, _P, c_int64, _P]),
    "runia_eigh_workspace_bytes": (c_size_t, [c_int]),
    "runia_eigh_f64": (c_int, [_P, c_int, _P, _P, _P, c_size_t, c_int, _P, _P]),
    "runia_cholesky_f64": (c_int, [_P, c_int, c_int, c_double, c_double, _P, _P, _P]),
    "runia_tril_inverse_f64": (c_int, [_P, c_int, c_int, _P, _P]),
    "runia_roi_align_f32": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int64, c_int, c_int, c_float, c_int, c_int,
                                    _P, _P]),
    "runia_roi_align_mean_f32": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int64, c_int, c_int, c_float, c_int,
                                         c_int, _P, _P, _P]),
}
EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(_lib, _name)  # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


def call(name, *args):
    """Calls an int-returning entry point and raises on a non-zero status."""
    rc = getattr(_lib, name)(*args)
    if rc != RUNIA_OK:
        msg = _lib.runia_b200_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(f"{name}: {msg}")
        if rc < 0:
            raise ValueError(f"{name}: {msg}")
        raise RuniaB200Error(f"{name}: CUDA error {rc}: {msg}")


def raw(name):
    return getattr(_lib, name)


def launch_count() -> int:
    return int(_lib.runia_b200_launch_count())
