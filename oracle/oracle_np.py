"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (NumPy/SciPy, float64 where the reference is float64) restatement of the reference's
post-hoc OoD scoring hot path.  Every function cites the reference lines it follows
(paths relative to /root/reference).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module, and only as the
checker or the timed CPU baseline -- never on the product path (`runia_core_b200/` never
imports `oracle`).

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against
(1) the reference's own hard-coded golden vectors (tests/unit_test_postprocessors.py,
unit_test_baselines.py, unit_test_feature_extraction.py, unit_test_dim_reduction.py,
unit_test_metrics.py) and (2) fixtures under tests/golden/ produced by importing the
unmodified reference source in the build container (oracle/gen_golden.py).
Exceptions (stated in DESIGN.md): GMM / DDU and ViM reference goldens are not reproducible
across LAPACK/torch versions even with the reference's own code; those
three are pinned only against fixtures generated here on well-conditioned data.

Third-party arithmetic restated here (not vendored in the reference):
* entropy-estimators==0.0.1 `continuous.get_h`          -> `get_h`
* faiss (faiss-gpu==1.7.2) `IndexFlatL2.search`          -> `flat_l2_search`
* torchmetrics==1.8.2 binary auroc / roc                 -> `binary_roc`, `auroc_fpr95`
* scikit-learn `EmpiricalCovariance`, `KernelDensity(gaussian)`, `PCA.transform`
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy import linalg as _sla
from scipy.special import digamma, logsumexp, softmax

FLT_MAX = float(np.finfo(np.float32).max)

# --------------------------------------------------------------------------------------
# (a1) MC-dropout latent-sample entropy.  evaluation/entropy.py:20-93
# --------------------------------------------------------------------------------------


def get_h(x, k=1, min_dist=1e-5):
    """Kozachenko-Leonenko kNN entropy, max-norm (entropy_estimators.continuous.get_h as
    called at evaluation/entropy.py:35,68,79-81 with norm="max", min_dist=1e-5).
    Brute-force distances instead of a cKDTree; same numbers."""
    x = np.asarray(x, np.float64)
    if x.ndim == 1:
        x = x[:, None]
    n, d = x.shape
    dist = np.abs(x[:, None, :] - x[None, :, :]).max(-1)  # Chebyshev, includes self (0)
    r = np.sort(dist, axis=1)[:, k]  # k-th neighbour, self excluded
    r = np.where(r < min_dist, min_dist, r)
    return -digamma(k) + digamma(n) + (d / float(n)) * np.sum(np.log(2.0 * r))


def entropy_k(n_mc: int) -> int:
    """evaluation/entropy.py:66"""
    return 5 if n_mc > 5 else n_mc - 1


def get_dl_h_z(z, n_mc, chunk=256):
    """evaluation/entropy.py:41-93, vectorised over items.  z: [N*n_mc, D] (item-major).
    Returns (h_mvn [N,1] f64, h_z [N,D] f64)."""
    z = np.asarray(z)
    n_items = z.shape[0] // n_mc
    D = z.shape[1]
    k = entropy_k(n_mc)
    z = z[: n_items * n_mc].reshape(n_items, n_mc, D)
    const = -digamma(k) + digamma(n_mc)
    h_mvn = np.empty((n_items, 1), np.float64)
    h_z = np.empty((n_items, D), np.float64)
    for s in range(0, n_items, chunk):
        zc = z[s:s + chunk].astype(np.float64)
        ad = np.abs(zc[:, :, None, :] - zc[:, None, :, :])  # [c, n, n, D]
        # per-dimension entropies (entropy.py:73-84): 1-D kNN distance per (item, dim)
        r = np.sort(ad, axis=2)[:, :, k, :]  # [c, n, D]
        r = np.where(r < 1e-5, 1e-5, r)
        h_z[s:s + chunk] = const + np.log(2.0 * r).sum(1) / n_mc
        # joint entropy (entropy.py:67-71): Chebyshev distance over all D dims
        cheb = ad.max(-1)  # [c, n, n]
        rj = np.sort(cheb, axis=2)[:, :, k]
        rj = np.where(rj < 1e-5, 1e-5, rj)
        h_mvn[s:s + chunk, 0] = const + (D / float(n_mc)) * np.log(2.0 * rj).sum(1)
    return h_mvn, h_z


def get_dl_h_z_faithful(z, n_mc):
    """Same numbers as `get_dl_h_z`, but with the reference's loop structure
    (entropy.py:56-84, parallel_run=False): one estimator call per item for the joint
    entropy and one per (item, dimension), each building a cKDTree like the third-party
    estimator does.  Used as the timed CPU baseline ("port")."""
    from scipy.spatial import cKDTree

    z = np.asarray(z)
    k = entropy_k(n_mc)
    items = np.split(z, z.shape[0] // n_mc)

    def _h(x):
        x = np.asarray(x, np.float64)
        if x.ndim == 1:
            x = x[:, None]
        n, d = x.shape
        r = cKDTree(x).query(x, k + 1, eps=0, p=np.inf)[0][:, -1]
        r[r < 1e-5] = 1e-5
        return -digamma(k) + digamma(n) + (d / float(n)) * np.sum(np.log(2 * r))

    h_mvn = np.expand_dims(np.array([_h(s) for s in items]), 1)
    h_z = np.asarray([[_h(s[:, j]) for j in range(s.shape[1])] for s in items])
    return h_mvn, h_z


# --------------------------------------------------------------------------------------
# (a2) PCA projection.  dimensionality_reduction.py:52-87 (sklearn PCA.transform)
# --------------------------------------------------------------------------------------


def pca_transform(X, mean, components, explained_variance, whiten=True):
    """sklearn _BasePCA.transform as reached from dimensionality_reduction.py:86:
    Z = X @ C^T - mean @ C^T, then Z /= sqrt(explained_variance) when whiten."""
    X = np.asarray(X)
    Z = X @ components.T - (mean.reshape(1, -1) @ components.T)
    if whiten:
        scale = np.sqrt(explained_variance)
        min_scale = np.finfo(scale.dtype).eps
        scale = np.where(scale < min_scale, min_scale, scale)
        Z = Z / scale
    return Z


# --------------------------------------------------------------------------------------
# (a3) LaREM Mahalanobis.  inference/postprocessors.py:202-244
# --------------------------------------------------------------------------------------


def empirical_precision(Xc, assume_centered=False):
    """sklearn EmpiricalCovariance(...).fit(Xc).precision_  (divisor N, pinvh)."""
    Xc = np.asarray(Xc)
    if assume_centered:  # np.dot(X.T, X) / n in the input dtype
        cov = (Xc.T @ Xc) / Xc.shape[0]
    else:  # np.cov(X.T, bias=1): promotes to float64, re-centres, divides by N
        Xd = Xc.astype(np.float64)
        Xd = Xd - Xd.mean(0)
        cov = (Xd.T @ Xd) / Xd.shape[0]
    return _sla.pinvh(cov, check_finite=False), cov


def md_fit(X):
    """postprocessors.py:212-220 -> (feats_mean [1,d], precision [d,d] f64)"""
    mean = np.mean(X, 0, keepdims=True)
    prec, _ = empirical_precision(X - mean)
    return mean, prec


def md_score(X, mean, precision, chunk=8192):
    """postprocessors.py:241-242, rowwise (the reference's N x N product restricted to its
    diagonal; identical numbers without the O(N^2) temporary)."""
    out = np.empty(X.shape[0], np.float64)
    for s in range(0, X.shape[0], chunk):
        diff = X[s:s + chunk] - mean
        out[s:s + chunk] = -np.einsum("ij,jk,ik->i", diff, precision, diff, optimize=True)
    return out


def md_score_faithful(X, mean, precision):
    """postprocessors.py:241-242 verbatim structure (materialises N x N); timed baseline."""
    diff = X - mean
    return -np.diag(np.matmul(np.matmul(diff, precision), np.transpose(diff)))


# --------------------------------------------------------------------------------------
# (a4) LaRED Gaussian KDE.  inference/postprocessors.py:109-128,150-178
# --------------------------------------------------------------------------------------


def kde_score(Q, bank, bandwidth=1.0, chunk=256):
    """sklearn KernelDensity(gaussian, bw).score_samples closed form (exact because
    atol=rtol=0): logsumexp_i(-|q-x_i|^2 / 2h^2) - log Nb - d/2 log(2 pi h^2)."""
    Q = np.asarray(Q, np.float64)
    B = np.asarray(bank, np.float64)
    nb, d = B.shape
    out = np.empty(Q.shape[0], np.float64)
    b2 = (B * B).sum(1)
    for s in range(0, Q.shape[0], chunk):
        q = Q[s:s + chunk]
        d2 = (q * q).sum(1)[:, None] + b2[None, :] - 2.0 * (q @ B.T)
        if nb * d * len(q) <= 50_000_000:  # exact differences when affordable
            d2 = ((q[:, None, :] - B[None, :, :]) ** 2).sum(-1)
        out[s:s + chunk] = logsumexp(-0.5 * d2 / bandwidth**2, axis=1)
    return out - np.log(nb) - 0.5 * d * np.log(2.0 * np.pi * bandwidth**2)


# --------------------------------------------------------------------------------------
# (a5) kNN.  inference/postprocessors.py:385-423, 825-883; inference/funcs.py:105-115
# --------------------------------------------------------------------------------------


def normalizer(x):
    """inference/funcs.py:115 verbatim semantics (dtype preserving)."""
    return x / (np.linalg.norm(x, ord=2, axis=-1, keepdims=True) + 1e-10)


def seq32_tree_sum(terms):
    """Fixed-order float64 sum shared bit-for-bit with the CUDA kernels (csrc/common.cuh
    `warp_tree_sum_f64`): term j goes to lane j % 32, every lane adds its terms in increasing j,
    then the 32 partial sums are combined by an xor butterfly (16, 8, 4, 2, 1).  terms: [..., d]."""
    t = np.asarray(terms, np.float64)
    d = t.shape[-1]
    pad = (-d) % 32
    if pad:
        t = np.concatenate([t, np.zeros(t.shape[:-1] + (pad,))], axis=-1)
    t = t.reshape(t.shape[:-1] + (-1, 32))
    p = np.zeros(t.shape[:-2] + (32,))
    for c in range(t.shape[-2]):  # sequential per lane
        p = p + t[..., c, :]
    for off in (16, 8, 4, 2, 1):
        p = p[..., :off] + p[..., off:2 * off]
    return p[..., 0]


def normalize_rows_exact(x):
    """Deterministic row normaliser shared bit-for-bit with the CUDA path
    (runia_normalize_rows): squared norm = seq32_tree_sum(float64(x)^2), n = sqrt(.),
    x_hat = float32(float64(x) / (n + 1e-10)).  Differs from `normalizer` (float32 NumPy,
    funcs.py:115) by at most 1 float32 ulp per element."""
    xd = np.ascontiguousarray(x).astype(np.float64)
    nrm = np.sqrt(seq32_tree_sum(xd * xd))
    return (xd / (nrm[:, None] + 1e-10)).astype(np.float32)


def flat_l2_search_tree(bank_f32, queries_f32, k):
    """faiss.IndexFlatL2.search restated with a total order: exact squared L2 between float32
    vectors evaluated in float64 as seq32_tree_sum((q - b)^2) -- differences of float32 values
    are exact in float64, each square is rounded once, the sum order is fixed -- neighbours
    sorted by (distance, index) ascending; distances returned as float32; FLT_MAX / -1 padding
    when k > ntotal (tests/unit_test_postprocessors.py:355-383)."""
    B = np.ascontiguousarray(bank_f32, np.float32).astype(np.float64)
    Q = np.ascontiguousarray(queries_f32, np.float32).astype(np.float64)
    nq, nb = Q.shape[0], B.shape[0]
    D = np.full((nq, k), FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    kk = min(k, nb)
    ar = np.arange(nb)
    for r in range(nq):
        diff = B - Q[r][None, :]
        acc = seq32_tree_sum(diff * diff)
        order = np.lexsort((ar, acc))[:kk]
        D[r, :kk] = acc[order].astype(np.float32)
        I[r, :kk] = order
    return D, I


flat_l2_search = flat_l2_search_tree


def knn_score(test, bank_normed_f32, k):
    """postprocessors.py:414-423 / 872-882: minus the squared distance to the k-th neighbour."""
    qn = normalize_rows_exact(test)
    D, I = flat_l2_search(bank_normed_f32, qn, k)
    return -D[:, -1], I


def knn_score_faithful(test, bank_normed_f32, k):
    """Loop structure of postprocessors.py:417-421 (one query per search call, float32 direct
    differences like faiss' non-BLAS path); timed CPU baseline."""
    out = []
    for feat in test:
        q = normalizer(feat.reshape(1, -1)).astype(np.float32)
        diff = bank_normed_f32 - q
        d2 = np.einsum("ij,ij->i", diff, diff)
        if k <= len(d2):
            kth = np.partition(d2, k - 1)[k - 1]
        else:
            kth = np.float32(FLT_MAX)
        out.append(-kth)
    return np.asarray(out, np.float32)


# --------------------------------------------------------------------------------------
# (a6) class-conditional Mahalanobis.  funcs.py:33-102; postprocessors.py:276-357
# --------------------------------------------------------------------------------------


def mahalanobis_fit(feats, labels, num_classes):
    """funcs.py:48-66 -> (class_mean [C,d], precision [d,d] f64)."""
    class_mean, centered = [], []
    for c in range(num_classes):
        xs = feats[labels == c]
        if len(xs) == 0:
            warnings.warn(f"No train examples for class {c}")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            class_mean.append(xs.mean(0))
        centered.append(xs - class_mean[c].reshape(1, -1))
    class_mean = np.stack(class_mean)
    prec, _ = empirical_precision(np.concatenate(centered).astype(np.float32))
    return class_mean, prec


def mahalanobis_score(feats, class_mean, precision, num_classes):
    """funcs.py:87-102: max over classes of -(x-mu_c)^T P (x-mu_c); NaN -> -inf."""
    scores = np.empty((feats.shape[0], num_classes), np.float64)
    for c in range(num_classes):
        t = feats - class_mean[c].reshape(1, -1)
        scores[:, c] = -np.einsum("ij,jk,ik->i", t, precision, t, optimize=True)
    scores[np.isnan(scores)] = -np.inf
    return scores.max(1)


# --------------------------------------------------------------------------------------
# (a7) ViM.  postprocessors.py:1021-1112
# --------------------------------------------------------------------------------------


def vim_fit(train_feats, train_logits, w, b):
    """postprocessors.py:1048-1080 -> (u, DIM, NS, alpha)"""
    u = -np.matmul(np.linalg.pinv(w), b)
    d = train_feats.shape[-1]
    DIM = 1000 if d >= 2048 else (512 if d >= 768 else d // 2)
    xc = train_feats - u
    cov = (xc.T @ xc) / xc.shape[0]  # EmpiricalCovariance(assume_centered=True)
    eig_vals, eig_vecs = np.linalg.eig(cov)
    NS = np.ascontiguousarray((eig_vecs.T[np.argsort(eig_vals * -1)[DIM:]]).T)
    vlogit = np.linalg.norm(np.matmul(xc, NS), axis=-1)
    alpha = train_logits.max(axis=-1).mean() / vlogit.mean()
    return u, DIM, NS, alpha


def vim_score(feats, logits, u, NS, alpha):
    """postprocessors.py:1105-1112 (no sign flip)."""
    vlogit = np.linalg.norm(np.matmul(feats - u, NS), axis=-1) * alpha
    return -vlogit + logsumexp(logits, axis=-1)


# --------------------------------------------------------------------------------------
# (a8) logit-space scores.  postprocessors.py:519-551, 580-608, 650-691; funcs.py:347-375
# --------------------------------------------------------------------------------------


def energy_score(logits):
    return logsumexp(logits, axis=1)


def msp_score(logits):
    return np.max(softmax(logits, axis=1), axis=1)


def generalized_entropy(probs, gamma, M):
    ps = np.sort(probs, axis=1)[:, -M:]
    return -np.sum(ps**gamma * (1 - ps) ** gamma, axis=1)


def gen_score(logits, gamma, M):
    return generalized_entropy(softmax(logits, axis=1), gamma, M)


# --------------------------------------------------------------------------------------
# (a9) DDU / GMM.  funcs.py:265-344; postprocessors.py:458-492, 731-786
# --------------------------------------------------------------------------------------


def gmm_fit_np(feats, labels, num_classes):
    """funcs.py:287-342 in float64 NumPy: per-class mean and covariance X^T X/(n-1), empty
    classes dropped, jitter ladder until the Cholesky factorisation succeeds.
    Returns (means [C',d], chol_lower [C',d,d], jitter)."""
    means, covs = [], []
    for c in range(num_classes):
        xs = np.asarray(feats[labels == c], np.float64)
        if len(xs) == 0:
            continue
        mu = xs.mean(0)
        n = xs.shape[0]
        n = n + 1 if n == 1 else n
        xc = xs - mu
        covs.append(xc.T @ xc / (n - 1))
        means.append(mu)
    means, covs = np.stack(means), np.stack(covs)
    for jitter in [0] + [10.0**e for e in range(-20, 0)]:
        try:
            L = np.linalg.cholesky(covs + jitter * np.eye(covs.shape[1])[None])
            if np.all(np.isfinite(L)):
                break
        except np.linalg.LinAlgError:
            continue
    return means, L, jitter


def gmm_lse_score(feats, means, chol):
    """logsumexp_c log N(x; mu_c, L_c L_c^T)  (postprocessors.py:490-491, 783-784)."""
    X = np.asarray(feats, np.float64)
    d = X.shape[1]
    lp = np.empty((X.shape[0], means.shape[0]), np.float64)
    for c in range(means.shape[0]):
        y = _sla.solve_triangular(chol[c], (X - means[c]).T, lower=True)
        lp[:, c] = (-0.5 * (y * y).sum(0) - np.log(np.diag(chol[c])).sum()
                    - 0.5 * d * np.log(2 * np.pi))
    return logsumexp(lp, axis=1)


# --------------------------------------------------------------------------------------
# (a10) ReAct / DICE / DICE+ReAct / ASH.  postprocessors.py:1152-1621; funcs.py:124-261
# --------------------------------------------------------------------------------------


def react_threshold(train_feats, percentile):
    return np.percentile(train_feats.flatten(), percentile)


def react_score(feats, w, b, thr):
    return logsumexp(np.matmul(feats.clip(max=thr), w.T) + b, axis=1)


def dice_masked_weight(train_feats, w, percentile):
    """funcs.py:171-180: contribution = mean_train * W, global percentile, strict >."""
    info = np.asarray(train_feats, np.float32).mean(0)
    contrib = info[None, :] * w
    thresh = np.percentile(contrib, percentile)
    return (w * (contrib > thresh)).astype(np.float32), thresh


def dice_score(feats, masked_w, b, clip=None):
    x = np.asarray(feats, np.float32)
    if clip is not None:
        x = x.clip(max=clip)
    return logsumexp((x[:, None, :] * masked_w[None]).sum(2) + b, axis=1)


def ash_s(x, percentile):
    """funcs.py:243-261"""
    s1 = x.sum(axis=1)
    n = x.shape[1]
    k = n - int(np.round(n * percentile / 100.0))
    idx = np.argpartition(x, -k)[:, -k:]
    top_k = np.partition(x, -k)[:, -k:]
    scattered = np.zeros_like(x)
    np.put_along_axis(scattered, indices=idx, values=top_k, axis=1)
    s2 = scattered.sum(axis=1)
    return scattered * np.exp((s1 / s2)[:, None])


def ash_score(feats, w, b, percentile):
    return logsumexp(np.matmul(ash_s(feats, percentile), w.T) + b, axis=1)


def ash_s_intended(x, percentile):
    """The rule funcs.py:243-261 describes ("keeps the top-k elements per row"), made deterministic: the k largest
    activations stay at their own positions (ties with the k-th value: lowest index first), the rest are zeroed,
    rows are scaled by exp(s1 / s2).  `ash_s` above is the LITERAL upstream code: it scatters np.partition's values
    to np.argpartition's indices, two calls that NumPy does not promise to order alike -- with NumPy >= 2 (SIMD
    quickselect for np.partition) they disagree on a few per cent of wide rows, i.e. the kept VALUES land permuted
    among the kept POSITIONS, and which rows depends on the NumPy build and the CPU.  scripts/ash_literal_delta.py
    measures what that does to the scores and to AUROC / FPR@95; the CUDA kernels implement this function."""
    x = np.asarray(x)
    n, d = x.shape
    k = d - int(np.round(d * percentile / 100.0))
    order = np.lexsort((np.broadcast_to(np.arange(d), x.shape), -x), axis=1)[:, :k]
    kept = np.zeros_like(x)
    np.put_along_axis(kept, order, np.take_along_axis(x, order, axis=1), axis=1)
    s1 = x.sum(axis=1)
    s2 = kept.sum(axis=1)
    with np.errstate(all="ignore"):
        return kept * np.exp((s1 / s2)[:, None])


def ash_score_intended(feats, w, b, percentile):
    return logsumexp(np.matmul(ash_s_intended(feats, percentile), w.T) + b, axis=1)


# --------------------------------------------------------------------------------------
# thresholds and the parity metric.  abstract_classes.py:408-424; evaluation/metrics.py:60-81
# --------------------------------------------------------------------------------------


def method_threshold(scores, z=1.645):
    return float(np.mean(scores)) - z * float(np.std(scores))


def binary_roc(scores, labels):
    """torchmetrics.functional.roc(task="binary") restated: sigmoid when any score is outside
    [0,1]; descending sort; one point per distinct score; (0,0) prepended."""
    s = np.asarray(scores).reshape(-1)
    y = np.asarray(labels).reshape(-1).astype(np.int64)
    if not ((s >= 0).all() and (s <= 1).all()):
        with np.errstate(over="ignore"):
            s = (1.0 / (1.0 + np.exp(-s.astype(s.dtype)))).astype(s.dtype)
    order = np.argsort(-s, kind="stable")
    s, y = s[order], y[order]
    distinct = np.nonzero(s[1:] - s[:-1])[0]
    idx = np.concatenate([distinct, [y.size - 1]])
    tps = np.cumsum(y)[idx]
    fps = 1 + idx - tps
    tps = np.concatenate([[0], tps])
    fps = np.concatenate([[0], fps])
    fpr = fps / fps[-1] if fps[-1] > 0 else np.zeros_like(fps, float)
    tpr = tps / tps[-1] if tps[-1] > 0 else np.zeros_like(tps, float)
    return fpr.astype(np.float64), tpr.astype(np.float64)


def binary_clf_curve(scores, labels):
    """torchmetrics `_binary_clf_curve` after `_binary_precision_recall_curve_format` (sigmoid when
    any score is outside [0, 1]): (fps, tps) at every distinct score, descending."""
    s = np.asarray(scores).reshape(-1)
    y = np.asarray(labels).reshape(-1).astype(np.int64)
    if not ((s >= 0).all() and (s <= 1).all()):
        with np.errstate(over="ignore"):
            s = (1.0 / (1.0 + np.exp(-s.astype(s.dtype)))).astype(s.dtype)
    order = np.argsort(-s, kind="stable")
    s, y = s[order], y[order]
    idx = np.concatenate([np.nonzero(s[1:] - s[:-1])[0], [y.size - 1]])
    tps = np.cumsum(y)[idx]
    return 1 + idx - tps, tps


def ood_metrics(ind_scores, ood_scores):
    """evaluation/metrics.py:60-81 (`get_auroc_results`): (auroc, fpr@95, aupr) with InD as the
    positive class.  AUPR = sklearn.metrics.auc over torchmetrics' precision_recall_curve (points
    reversed, (recall 0, precision 1) appended)."""
    ind = np.asarray(ind_scores).reshape(-1)
    ood = np.asarray(ood_scores).reshape(-1)
    scores = np.concatenate([ind, ood])
    labels = np.concatenate([np.ones(ind.size, np.int64), np.zeros(ood.size, np.int64)])
    fps, tps = binary_clf_curve(scores, labels)
    fpr = np.concatenate([[0.0], fps / fps[-1]])
    tpr = np.concatenate([[0.0], tps / tps[-1]])
    trapz = np.trapezoid if hasattr(np, "trapezoid") else np.trapz
    auroc = float(trapz(tpr, fpr))
    fpr95 = float(fpr[np.where(tpr >= 0.95)[0][0]])
    precision = np.concatenate([(tps / (tps + fps))[::-1], [1.0]])
    recall = np.concatenate([(tps / tps[-1])[::-1], [0.0]])
    aupr = float(-trapz(precision, recall))  # recall decreases: sklearn's auc flips the sign
    return auroc, fpr95, aupr


def auroc_fpr95(ind_scores, ood_scores):
    """evaluation/metrics.py:60-76: InD is the positive class."""
    ind = np.asarray(ind_scores).reshape(-1)
    ood = np.asarray(ood_scores).reshape(-1)
    scores = np.concatenate([ind, ood])
    labels = np.concatenate([np.ones(ind.size, np.int64), np.zeros(ood.size, np.int64)])
    fpr, tpr = binary_roc(scores, labels)
    auroc = float(np.trapezoid(tpr, fpr)) if hasattr(np, "trapezoid") else float(np.trapz(tpr, fpr))
    fpr95 = float(fpr[np.where(tpr >= 0.95)[0][0]])
    return auroc, fpr95


# --------------------------------------------------------------------------------------
# (f4) EigenScore.  llm_uncertainty/scores.py:49-66
# --------------------------------------------------------------------------------------
def eigen_score_faithful(E, alpha=1e-3):
    """The reference's own route: d x d covariance of the float32 embeddings (torch.cov keeps
    float32), cast to float64, full SVD of cov + alpha I, mean log singular value."""
    E = np.asarray(E, np.float32)
    Ec = E - E.mean(0, keepdims=True, dtype=np.float32)
    cov = (Ec.T @ Ec / np.float32(E.shape[0] - 1)).astype(np.float64)
    sv = np.linalg.svd(cov + alpha * np.eye(cov.shape[0]), compute_uv=False)
    return float(np.mean(np.log(sv)))


def eigen_score(E, alpha=1e-3):
    """Same quantity from the n x n Gram matrix of the centred samples in float64: the d x d covariance has
    rank <= n - 1 and shares its non-zero eigenvalues with the Gram matrix; the other d - n singular values
    of cov + alpha I are alpha."""
    E = np.asarray(E, np.float64)
    n, d = E.shape
    Ec = E - E.mean(0, keepdims=True)
    lam = np.linalg.eigvalsh(Ec @ Ec.T / (n - 1))
    return float((np.log(np.maximum(lam, 0.0) + alpha).sum() + (d - n) * np.log(alpha)) / d)


# --------------------------------------------------------------------------------------
# (f4) predictive entropy / mutual information.  inference/funcs.py:430-465
# --------------------------------------------------------------------------------------
def predictive_uncertainty(logits, n_mc):
    """float32 like the torch original: softmax per row, mean over the n_mc rows of an item,
    pred_h = -sum q log q, mi = pred_h - mean_s(-sum p log p)."""
    x = np.asarray(logits, np.float32)
    p = softmax(x, axis=1).astype(np.float32).reshape(-1, n_mc, x.shape[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        q = p.mean(1, dtype=np.float32)
        pred_h = -(q * np.log(q)).sum(1, dtype=np.float32)
        exp_h = (-(p * np.log(p)).sum(-1, dtype=np.float32)).mean(1, dtype=np.float32)
    return pred_h, pred_h - exp_h
