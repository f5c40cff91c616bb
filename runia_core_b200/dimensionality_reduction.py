"""PCA projection with the reference's signatures (`runia_core/dimensionality_reduction.py:26-87`).

The default fit stays on scikit-learn (randomized SVD consuming the global NumPy RNG stream exactly like
upstream, so seeded results are identical); the returned estimator is a `PCA` subclass whose
`transform` runs the projection GEMM on the GPU (`runia_pca_transform_*`).  `svd_solver="covariance_eigh"` (one of
scikit-learn's own solvers: exact eigenvectors of the covariance, no RNG) is fitted ON THE DEVICE for float32 data:
NumPy-ordered column means, float64 Gram matrix of the residuals, Jacobi eigendecomposition (`runia_eigh_f64`), with
sklearn's sign convention and derived attributes.  Parity definition for that fit: components, explained variances and
transformed rows equal scikit-learn's `PCA(svd_solver="covariance_eigh")` on the same float32 data promoted to float64
to 1e-5 (the rounding of the float32 centring).  The PaCMAP plotting helpers of that file are out of scope
(SURVEY.md section 2 #15)."""
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

    def _device_fit_ok(self, X) -> bool:
        return self.svd_solver == "covariance_eigh" and isinstance(self.n_components, (int, np.integer)) and \
            isinstance(X, np.ndarray) and X.dtype == np.float32 and X.ndim == 2 and \
            1 <= self.n_components <= min(X.shape) and X.shape[0] >= 2 and X.shape[1] <= 8192

    def _fit_device(self, X):
        """sklearn's `_fit_full(..., "covariance_eigh")` with the O(N d^2) and O(d^3) parts on the device."""
        n, d = X.shape
        k = int(self.n_components)
        means, _, xf, _ = _ops.class_means(X, None, 1)
        G, cs = _ops.centered_gram(xf, None, means)
        cov = _ops.covariance_from_gram(G, cs, n) * (n / (n - 1.0))        # np.cov(..., bias=1) -> divisor n - 1
        lam, V = _ops.eigh(cov)
        lam, V = lam[::-1].copy(), V[:, ::-1]
        lam[lam < 0.0] = 0.0                                              # rounding residue, like sklearn
        Vt = np.ascontiguousarray(V.T)
        signs = np.sign(Vt[np.arange(d), np.argmax(np.abs(Vt), axis=1)])  # svd_flip(u_based_decision=False)
        signs[signs == 0] = 1.0
        Vt *= signs[:, None]
        self.mean_ = to_host(means).reshape(-1).astype(np.float64)
        self.n_samples_, self.n_features_in_, self.n_components_ = n, d, k
        self.components_ = Vt[:k]
        self.explained_variance_ = lam[:k]
        total = lam.sum()
        self.explained_variance_ratio_ = lam[:k] / total if total > 0 else np.zeros(k)
        self.singular_values_ = np.sqrt(lam[:k] * (n - 1))
        self.noise_variance_ = float(lam[k:].mean()) if k < min(n, d) else 0.0
        self._b200_state = None
        return self

    def fit(self, X, y=None):
        self._b200_state = None
        if self._device_fit_ok(X):
            return self._fit_device(X)
        return super().fit(X, y)

    def fit_transform(self, X, y=None):
        self._b200_state = None
        if self._device_fit_ok(X):
            return self._fit_device(X).transform(X)
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
