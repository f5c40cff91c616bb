"""PCA projection with the reference's signatures (`runia_core/dimensionality_reduction.py:26-87`).

The fit stays on scikit-learn (randomized SVD consuming the global NumPy RNG stream exactly like
upstream, so seeded results are identical); the returned estimator is a `PCA` subclass whose
`transform` runs the projection GEMM on the GPU (`runia_pca_transform_f32`).  The PaCMAP plotting
helpers of that file are out of scope."""
from typing import Tuple

import numpy as np
import torch
from sklearn.decomposition import PCA

from . import _ops
from ._device import to_device, to_host

__all__ = ["apply_pca_ds", "apply_pca_ds_split", "apply_pca_transform", "B200PCA"]


class B200PCA(PCA):
    """sklearn PCA whose `transform` is the CUDA projection; `components_`, `mean_`,
    `explained_variance_` etc. are the fitted sklearn attributes."""

    _b200_state = None

    def _state(self):
        if self._b200_state is None:
            self._b200_state = _ops.pca_prepare(self.mean_, self.components_, self.explained_variance_, self.whiten)
        return self._b200_state

    def fit(self, X, y=None):
        self._b200_state = None
        return super().fit(X, y)

    def fit_transform(self, X, y=None):
        self._b200_state = None
        return super().fit_transform(X, y)

    def transform_device(self, X) -> torch.Tensor:
        """[N, D0] (host or device) -> [N, d] float32 CUDA tensor."""
        return _ops.pca_transform(X, self._state())

    def transform(self, X):
        in_dtype = X.dtype if isinstance(X, np.ndarray) else None
        out = to_host(self.transform_device(X))
        return out.astype(np.float64) if in_dtype == np.float64 else out


def apply_pca_ds(train_samples: np.ndarray, test_samples: np.ndarray, nro_components: int = 16,
                 svd_solver: str = "randomized", whiten: bool = True):
    """dimensionality_reduction.py:26-49"""
    pca = B200PCA(n_components=nro_components, svd_solver=svd_solver, whiten=whiten)
    train_ds = pca.fit_transform(train_samples)
    test_ds = pca.transform(test_samples)
    return train_ds, test_ds, pca


def apply_pca_ds_split(samples: np.ndarray, nro_components: int = 16, svd_solver: str = "randomized",
                       whiten: bool = True) -> Tuple[np.ndarray, PCA]:
    """Fits PCA on `samples`, returns (reduced samples, fitted estimator) --
    dimensionality_reduction.py:52-72."""
    pca = B200PCA(n_components=nro_components, svd_solver=svd_solver, whiten=whiten)
    return pca.fit_transform(samples), pca


def apply_pca_transform(samples: np.ndarray, pca_transform: PCA) -> np.ndarray:
    """dimensionality_reduction.py:75-87.  Accepts any fitted sklearn PCA."""
    if not isinstance(pca_transform, B200PCA):
        st = _ops.pca_prepare(pca_transform.mean_, pca_transform.components_, pca_transform.explained_variance_,
                              pca_transform.whiten)
        out = to_host(_ops.pca_transform(samples, st))
        return out.astype(np.float64) if getattr(samples, "dtype", None) == np.float64 else out
    return pca_transform.transform(samples)
