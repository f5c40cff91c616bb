"""Scoring helpers with the names and argument meaning of the reference's
`runia_core/inference/funcs.py` (cited per function).  Fits run once on the host exactly like
the reference (`setup()` is not the hot path); everything that
scores rows goes through the CUDA library."""
import warnings
from typing import Dict, Tuple, Union

import numpy as np
import torch
from scipy.linalg import pinvh
from sklearn.covariance import EmpiricalCovariance

from .. import _ops
from .._device import to_device, to_host

__all__ = [
    "RouteDICE",
    "ash_s_linear_layer",
    "gmm_fit",
    "generalized_entropy",
    "get_predictive_uncertainty_score",
    "mahalanobis_preprocess",
    "mahalanobis_postprocess",
    "normalizer",
]


def mahalanobis_preprocess(ind_data: Dict[str, np.ndarray], num_classes: int) -> Tuple[np.ndarray, np.ndarray]:
    """Class means [C, d] and the shared precision of the class-centred training features
    (funcs.py:33-66).  Classes without samples warn and get a NaN mean."""
    feats, labels = ind_data["train features"], ind_data["train labels"]
    if isinstance(feats, np.ndarray) and feats.dtype == np.float32 and feats.ndim == 2 and feats.shape[0] > 0:
        # device statistics (csrc/fit.cu): NumPy-ordered class means, float64 Gram matrix of the residuals
        means, counts, xf, lab = _ops.class_means(feats, labels, num_classes)
        for c in np.flatnonzero(counts == 0):
            warnings.warn(f"No train examples for class {c}")
        cov = _ops.centered_covariance(xf, lab, means, int(counts.sum()))
        if not np.isfinite(cov).all():
            raise ValueError("Input X contains NaN or infinity.")
        return to_host(means), (_ops.pinvh(cov) if cov.shape[0] >= 64 else pinvh(cov, check_finite=False))
    class_mean, centered = [], []
    for c in range(num_classes):
        xs = feats[labels == c]
        if len(xs) == 0:
            warnings.warn(f"No train examples for class {c}")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            class_mean.append(xs.mean(0))
        centered.append(xs - class_mean[c].reshape(1, -1))
    class_mean = np.stack(class_mean)
    ec = EmpiricalCovariance(assume_centered=False)
    ec.fit(np.concatenate(centered).astype(np.float32))
    return class_mean, ec.precision_


def mahalanobis_postprocess(feats: np.ndarray, class_mean: np.ndarray, precision: np.ndarray,
                            num_classes: int, _state=None) -> np.ndarray:
    """max_c -(x - mu_c)^T P (x - mu_c), float64 (funcs.py:69-102).  CUDA: one contraction against
    the factored precision + per-class squared distances (runia_classcond_mahalanobis_f32)."""
    st = _state if _state is not None else _ops.classcond_prepare(class_mean[:num_classes], precision)
    return to_host(_ops.classcond_score(feats, st, torch.float64))


def normalizer(x):
    """x / (||x||_2 + 1e-10) along the last axis (funcs.py:105-115), float32 result."""
    a = np.asarray(x) if not isinstance(x, torch.Tensor) else x
    shape = tuple(a.shape)
    out = to_host(_ops.normalize_rows(a.reshape(-1, shape[-1])))
    return out.reshape(shape)


class RouteDICE(torch.nn.Linear):
    """DICE sparsified linear layer (funcs.py:124-190): weights whose contribution
    mean_train[d] * W[c, d] is not above the global p-th percentile are masked out.
    `forward` returns the logits x @ masked_W^T + b; the OoD score path of the DICE
    postprocessors fuses clip + this product + log-sum-exp into one kernel instead."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, p: int = 90,
                 conv1x1: bool = False, info: Union[None, np.ndarray] = None):
        assert 0 < p < 100, "p must be greater than 0 and less than 100"
        if info is not None:
            assert isinstance(info, np.ndarray), "info must be a numpy array or None"
        super().__init__(in_features, out_features, bias)
        if conv1x1:
            self.weight = torch.nn.Parameter(torch.Tensor(out_features, in_features, 1, 1))
        self.p = p
        self.info = info
        self.masked_w = None
        self.contrib = None
        self.thresh = None

    def calculate_mask_weight(self):
        w = self.weight.data.cpu().numpy()
        self.contrib = self.info[None, :] * w
        self.thresh = np.percentile(self.contrib, self.p)
        mask = torch.Tensor((self.contrib > self.thresh))
        self.masked_w = to_device((self.weight.data.squeeze().cpu() * mask).contiguous(), torch.float32)

    def forward(self, x):
        if self.masked_w is None:
            self.calculate_mask_weight()
        bias = None if self.bias is None else to_device(self.bias.detach(), torch.float32)
        return _ops.linear(x, self.masked_w, bias)


def ash_s_linear_layer(x: np.ndarray, percentile: int = 85):
    """ASH-S pruning + rescaling of a feature matrix (funcs.py:230-261).  Returns the shaped
    features; the ASH postprocessor itself uses the fused kernel and never materialises them."""
    assert x.ndim == 2
    assert 0 <= percentile <= 100
    n = x.shape[1]
    k = n - int(np.round(n * percentile / 100.0))
    return to_host(_ops.ash_prune(x, k))


def _gmm_fit_device(embeddings: torch.Tensor, labels: torch.Tensor, num_classes: int):
    """gmm_fit on the device kernels: NumPy-ordered class means (`runia_class_mean_f32`), the float64 Gram matrix of
    each class's float32 residuals (`runia_centered_gram_f64`; upstream multiplies them in float32), covariance
    G / (n - 1), and the jitter ladder on a float64 Cholesky (`runia_cholesky_f64`) that gives up where a float32
    factorisation would (pivot below 6e-8 of its diagonal entry).  Parity definition: with well-conditioned classes
    (n_c >> d) the mixture's log-densities agree with the reference's float32 torch fit to 1e-5 relative (fixtures);
    rank-deficient classes are decided by rounding in the reference itself."""
    jitters = [0] + [10**e for e in range(-20, 0, 1)]
    means, counts, xf, lab = _ops.class_means(embeddings, labels, num_classes)
    keep = [c for c in range(num_classes) if counts[c] > 0]
    d = xf.shape[1]
    covs = torch.empty((len(keep), d, d), dtype=torch.float64, device=xf.device)
    for i, c in enumerate(keep):
        lab_c = torch.where(lab == c, 0, -1).to(torch.int32)
        G, _ = _ops.centered_gram(xf, lab_c, means[c:c + 1].contiguous())
        n = int(counts[c])
        n = n + 1 if n == 1 else n
        covs[i] = G / (n - 1)
    loc = means[keep].contiguous()
    gmm, jitter_eps = None, None
    for jitter_eps in jitters:
        L, fail = _ops.cholesky_batch(covs, jitter=float(jitter_eps), rel_pivot=6e-8)
        if (fail == 0).all() and bool(torch.isfinite(L).all()):
            gmm = torch.distributions.MultivariateNormal(loc=loc, scale_tril=L.to(torch.float32))
            break
    return gmm, jitter_eps


def gmm_fit(embeddings: torch.Tensor, labels: torch.Tensor, num_classes: int):
    """Class-wise Gaussian mixture (funcs.py:265-344): mean and covariance X^T X / (n-1) per class,
    empty classes dropped, smallest jitter of [0, 1e-20 .. 1e-1] for which the Cholesky
    factorisation exists.  Returns (MultivariateNormal, jitter).  float32 embeddings are fitted by the device kernels
    (`_gmm_fit_device`); other dtypes keep the reference's torch expression."""
    if torch.cuda.is_available() and isinstance(embeddings, torch.Tensor) and embeddings.dtype == torch.float32 and \
            embeddings.dim() == 2 and embeddings.shape[0] > 0 and embeddings.shape[1] <= 2048:
        return _gmm_fit_device(embeddings.detach(), labels, num_classes)
    jitters = [0] + [10**e for e in range(-20, 0, 1)]
    with torch.no_grad():
        means, covs = [], []
        for c in range(num_classes):
            xs = embeddings[labels == c]
            mu = torch.mean(xs, dim=0)
            n = xs.shape[0]
            n = n + 1 if n == 1 else n
            xc = xs - mu
            means.append(mu)
            covs.append(xc.t().mm(xc) / (n - 1))
        means, covs = torch.stack(means), torch.stack(covs)
        keep = ~torch.any(means.isnan(), dim=1)
        if not bool(keep.all()):
            means, covs = means[keep], covs[keep]
        gmm, jitter_eps = None, None
        for jitter_eps in jitters:
            try:
                jitter = jitter_eps * torch.eye(covs.shape[1], device=covs.device).unsqueeze(0)
                gmm = torch.distributions.MultivariateNormal(loc=means, covariance_matrix=(covs + jitter))
            except RuntimeError as e:
                if "cholesky" in str(e):
                    continue
            except ValueError as e:
                if "found invalid values" in str(e):
                    continue
            break
    return gmm, jitter_eps


def generalized_entropy(probs, gamma, M):
    """-sum over the M largest entries of each row of p^gamma (1-p)^gamma (funcs.py:347-375).  The rows are used as
    they are (no re-normalisation), like upstream; float64 input is scored in float32 and widened."""
    p = probs.detach().cpu().numpy() if isinstance(probs, torch.Tensor) else np.asarray(probs)
    g = to_host(_ops.gen_entropy_from_probs(p, gamma, M))
    return g.astype(p.dtype if p.dtype.kind == "f" else np.float32)


def get_predictive_uncertainty_score(input_samples: torch.Tensor, mcd_nro_samples: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Predictive entropy and mutual information of MC-dropout logits (funcs.py:430-465):
    input_samples [N * mcd_nro_samples, C] -> (pred_h [N], mi [N]) float32 tensors on the input's device.
    One fused kernel (softmax, mean over the MC samples, both entropies) instead of five passes."""
    assert input_samples.shape[0] % mcd_nro_samples == 0, (
        "Input tensor first dimension must be " "divisible by the mcd_nro_samples"
    )
    from .. import _lib
    from .._device import stream_ptr

    x = to_device(input_samples, torch.float32)
    n_items, C = x.shape[0] // int(mcd_nro_samples), x.shape[1]
    pred_h = torch.empty((n_items,), dtype=torch.float32, device=x.device)
    mi = torch.empty((n_items,), dtype=torch.float32, device=x.device)
    _lib.call("runia_pred_uncertainty_f32", x.data_ptr(), n_items, int(mcd_nro_samples), C, pred_h.data_ptr(),
              mi.data_ptr(), stream_ptr())
    if isinstance(input_samples, torch.Tensor) and not input_samples.is_cuda:
        return pred_h.cpu(), mi.cpu()
    return pred_h, mi
