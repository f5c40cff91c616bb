"""(f2) setup() statistics on the device (csrc/fit.cu) against NumPy / sklearn -- the libraries the reference's
setup() calls (postprocessors.py:202-226, 283-318; funcs.py:33-66)."""
import warnings

import numpy as np
import pytest
from sklearn.covariance import EmpiricalCovariance

pytestmark = pytest.mark.gpu


def _data(n, d, C, seed, empty=None):
    rng = np.random.RandomState(seed)
    labels = rng.randint(0, C, n)
    if empty is not None:
        labels[labels == empty] = (empty + 1) % C
    mu = rng.standard_normal((C, d)).astype(np.float32)
    x = (mu[labels] + rng.standard_normal((n, d))).astype(np.float32)
    return x, labels


@pytest.mark.parametrize("n,d,C,empty", [(50_000, 512, 10, None), (3001, 37, 4, 2), (5, 3, 1, None), (70_000, 256, 1, None)])
def test_class_means_bit_identical_to_numpy(n, d, C, empty):
    from runia_core_b200 import _ops

    x, labels = _data(n, d, C, 1, empty)
    means, counts, _, _ = _ops.class_means(x, labels if C > 1 else None, C)
    means = means.cpu().numpy()
    for c in range(C):
        xs = x[labels == c] if C > 1 else x
        assert counts[c] == len(xs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            ref = xs.mean(0)
        assert np.array_equal(means[c], ref, equal_nan=True), (c, np.abs(means[c] - ref).max())
    if C == 1:
        assert np.array_equal(means, np.mean(x, 0, keepdims=True))


@pytest.mark.parametrize("n,d,C", [(50_000, 512, 10), (3001, 37, 4), (20, 130, 1), (200_000, 64, 3)])
def test_centered_covariance_matches_np_cov(n, d, C):
    from runia_core_b200 import _ops

    x, labels = _data(n, d, C, 2)
    labels[::17] = C + 3  # rows of an unused label are left out, like the reference's per-class gather
    means, counts, xf, lab = _ops.class_means(x, labels, C)
    n_used = int(counts.sum())
    cov = _ops.centered_covariance(xf, lab, means, n_used)
    keep = labels < C
    resid = x[keep] - means.cpu().numpy()[labels[keep]]
    ref = np.cov(resid.T, bias=1)
    assert n_used == keep.sum() and cov.dtype == np.float64
    assert np.abs(cov - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.array_equal(cov, cov.T)
    assert np.array_equal(cov, _ops.centered_covariance(xf, lab, means, n_used))  # fixed-order reduction


def test_md_setup_device_fit_matches_host_fit():
    from runia_core_b200.inference import postprocessors_dict

    rng = np.random.RandomState(5)
    d = 256
    A = np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
    train = (0.5 + rng.standard_normal((50_000, d)) @ A).astype(np.float32)
    test = (rng.standard_normal((4000, d)) @ A).astype(np.float32)
    md = postprocessors_dict["MD"]()
    md.setup(train)
    mean = np.mean(train, 0, keepdims=True)
    ec = EmpiricalCovariance(assume_centered=False).fit(train - mean)
    assert np.array_equal(md.feats_mean, mean) and md.feats_mean.shape == (1, d)
    assert np.array_equal(md.centered_data, train - mean)
    assert np.abs(md.precision - ec.precision_).max() <= 1e-9 * np.abs(ec.precision_).max()
    diff = (test - mean).astype(np.float64)
    ref = -np.einsum("ij,jk,ik->i", diff, ec.precision_, diff)
    np.testing.assert_allclose(md.postprocess(test), ref, rtol=1e-5)  # north-star tolerance is 1e-4
    with pytest.warns(UserWarning, match="already trained"):
        md.setup(train)


def test_mahalanobis_and_cmd_setup_device_fit():
    from runia_core_b200.inference import postprocessors_dict
    from runia_core_b200.inference.funcs import mahalanobis_preprocess

    C, d = 10, 128
    x, labels = _data(20_000, d, C, 7, empty=4)
    with pytest.warns(UserWarning, match="No train examples for class 4"):
        cm, prec = mahalanobis_preprocess({"train features": x, "train labels": labels}, C)
    with pytest.warns(UserWarning, match="No train examples for class 4"):  # float64 input: the host route
        cm64, prec64 = mahalanobis_preprocess({"train features": x.astype(np.float64), "train labels": labels}, C)
    assert np.isnan(cm[4]).all() and cm.dtype == np.float32
    ok = np.arange(C) != 4
    np.testing.assert_allclose(cm[ok], cm64[ok], rtol=0, atol=5e-5)  # float32 row-by-row sums vs float64
    # host fit on the same float32 residuals
    resid = np.concatenate([x[labels == c] - x[labels == c].mean(0) for c in range(C) if c != 4])
    ec = EmpiricalCovariance(assume_centered=False).fit(resid)
    assert np.abs(prec - ec.precision_).max() <= 1e-9 * np.abs(ec.precision_).max()

    cmd = postprocessors_dict["cMD"]()
    with pytest.warns(UserWarning, match="No examples for class 4"):
        cmd.setup(x, ind_train_labels=labels)
    assert np.array_equal(cmd.class_mean.numpy()[ok], cm[ok])
    test = x[:3000] + 0.3
    got = cmd.postprocess(test, pred_labels=None)
    diff = test[:, None, :].astype(np.float64) - cm[None, ok].astype(np.float64)
    ref = (-np.einsum("ncj,jk,nck->nc", diff, ec.precision_, diff)).max(1)
    np.testing.assert_allclose(got, ref, rtol=1e-4)


def test_fit_argument_contract():
    import torch

    from runia_core_b200 import _lib

    x = torch.zeros(8, 4, device="cuda")
    out = torch.zeros(4, device="cuda")
    assert _lib.raw("runia_class_mean_f32")(x.data_ptr(), None, 8, 4, 2, out.data_ptr(), None, None) < 0  # C > 1 without labels
    assert b"labels" in _lib.raw("runia_b200_last_error")()
    G = torch.zeros(4, 4, dtype=torch.float64, device="cuda")
    assert _lib.raw("runia_centered_gram_f64")(x.data_ptr(), None, None, 8, 4, 1, G.data_ptr(), None, out.data_ptr(), 8, None) < 0
    assert b"workspace" in _lib.raw("runia_b200_last_error")()


@pytest.mark.parametrize("n", [5, 64, 257, 512])
def test_device_eigh_pinvh_cholesky(n):
    """csrc/eigh.cu: the float64 Jacobi eigendecomposition behind pinvh(covariance) (sklearn EmpiricalCovariance ->
    scipy.linalg.pinvh, postprocessors.py:212-220), and the batched Cholesky behind gmm_fit (funcs.py:296-342), against
    NumPy / SciPy: eigenvalues to 1e-12 of |A|, A V = V diag(lambda) to 1e-12, V orthonormal, pinvh to 1e-9 (rank-deficient
    covariance with scipy's cut-off, an indefinite matrix with +x / -x eigenvalue pairs)."""
    from scipy.linalg import pinvh

    from runia_core_b200 import _ops

    rng = np.random.RandomState(n)
    X = rng.randn(3 * n, n) @ (np.eye(n) + 0.3 * rng.randn(n, n) / np.sqrt(n))
    cov = np.cov(X.T, bias=True)
    lam, V = _ops.eigh(cov)
    ref = np.linalg.eigvalsh(cov)
    assert np.abs(lam - ref).max() < 1e-12 * np.abs(ref).max()
    assert np.abs(cov @ V - V * lam).max() < 1e-12 * np.abs(cov).max() * n
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    P = _ops.pinvh(cov)
    assert np.abs(P - pinvh(cov)).max() < 1e-9 * np.abs(P).max()
    # rank-deficient: fewer rows than columns -> scipy's cut-off drops the null space
    Xr = rng.randn(max(2, n // 2), n)
    cr = np.cov(Xr.T, bias=True)
    Pr, Sr = _ops.pinvh(cr), pinvh(cr)
    assert np.abs(Pr - Sr).max() < 1e-8 * np.abs(Sr).max()
    # indefinite with +x / -x pairs (Q diag(+-) Q^T): the residual check re-decomposes it with a shift
    Q = np.linalg.qr(rng.randn(n, n))[0]
    d = np.concatenate([np.linspace(1, 2, n // 2), -np.linspace(1, 2, n - n // 2)])
    Ai = (Q * d) @ Q.T
    li, Vi = _ops.eigh(Ai)
    assert np.abs(li - np.sort(d)).max() < 1e-10 and np.abs(Ai @ Vi - Vi * li).max() < 1e-10
    # batched Cholesky with the failure flag
    mats = np.stack([cov + 0.1 * np.eye(n), cr, Ai])
    L, fail = _ops.cholesky_batch(mats)
    assert fail[1] > 0 or n <= 5  # rank-deficient
    L = L.cpu().numpy()
    assert fail[0] == 0 and np.abs(L[0] - np.linalg.cholesky(mats[0])).max() < 1e-10 * np.abs(L[0]).max()
    assert fail[2] > 0  # indefinite
    L2, fail2 = _ops.cholesky_batch(mats[1:2], jitter=1e-3)
    assert fail2[0] == 0 and np.abs(L2[0].cpu().numpy() - np.linalg.cholesky(cr + 1e-3 * np.eye(n))).max() < 1e-9


def test_pca_covariance_eigh_fit_on_device():
    """apply_pca_ds_split(..., svd_solver="covariance_eigh") (dimensionality_reduction.py:52-72 passes the solver name
    to sklearn) fitted on the device == sklearn's own covariance_eigh solver on the same data in float64."""
    from sklearn.decomposition import PCA

    import runia_core_b200 as R

    rng = np.random.RandomState(8)
    X = (0.5 + rng.randn(20_000, 96) @ (np.eye(96) + 0.4 * rng.randn(96, 96) / 10)).astype(np.float32)
    Z, pca = R.apply_pca_ds_split(X, nro_components=24, svd_solver="covariance_eigh")
    ref = PCA(n_components=24, svd_solver="covariance_eigh", whiten=True).fit(X.astype(np.float64))
    np.testing.assert_allclose(pca.explained_variance_, ref.explained_variance_, rtol=1e-6)
    np.testing.assert_allclose(pca.components_, ref.components_, atol=2e-6)
    np.testing.assert_allclose(pca.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-6)
    np.testing.assert_allclose(pca.singular_values_, ref.singular_values_, rtol=1e-6)
    assert abs(pca.noise_variance_ - ref.noise_variance_) < 1e-6 * ref.noise_variance_
    np.testing.assert_allclose(Z, ref.transform(X.astype(np.float64)), atol=2e-5)
    assert Z.dtype == np.float32 and pca.n_components_ == 24 and pca.n_samples_ == 20_000
    np.testing.assert_allclose(R.apply_pca_transform(X[:100], pca), Z[:100], atol=1e-6)


def test_gmm_fit_on_device_matches_torch_expression():
    """gmm_fit (funcs.py:265-344) on the device kernels vs the reference's float32 torch expression on the same data:
    means, Cholesky factors and DDU log-densities (well-conditioned classes; an empty class is dropped)."""
    import torch

    from runia_core_b200 import _ops
    from runia_core_b200.inference.funcs import gmm_fit

    rng = np.random.RandomState(9)
    C, d, n = 6, 48, 6000
    y = rng.randint(0, C, n)
    y[y == 4] = 5  # class 4 empty
    mu = rng.randn(C, d)
    x = (mu[y] + rng.randn(n, d) @ (np.eye(d) + 0.2 * rng.randn(d, d) / 7)).astype(np.float32)
    gmm, jit = gmm_fit(torch.from_numpy(x), torch.from_numpy(y.astype(np.float32)), C)
    assert jit == 0 and gmm.loc.shape == (C - 1, d) and gmm.loc.is_cuda
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    means = torch.stack([xt[yt == c].mean(0) for c in range(C) if (yt == c).any()])
    covs = torch.stack([torch.cov(xt[yt == c].T.double()) for c in range(C) if (yt == c).any()])
    ref = torch.distributions.MultivariateNormal(means.double(), covariance_matrix=covs)
    np.testing.assert_allclose(gmm.loc.cpu().numpy(), means.numpy(), atol=2e-6)
    np.testing.assert_allclose(gmm.scale_tril.cpu().numpy(), ref.scale_tril.numpy(), rtol=1e-4, atol=1e-5)
    st = _ops.gmm_prepare(gmm.loc.cpu().numpy(), gmm.scale_tril.cpu().numpy())
    got = _ops.gmm_lse(x[:500], st).cpu().numpy()
    want = torch.logsumexp(ref.log_prob(xt[:500, None, :].double()), dim=1).numpy()
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-4


@pytest.mark.parametrize("n,d", [(50_000, 512), (3001, 70), (9, 5)])
def test_shifted_gram_matches_float64_numpy(n, d):
    """ViM.setup's covariance: EmpiricalCovariance(assume_centered=True).fit(train - u) with a float64 u
    (postprocessors.py:1060-1064) = (x - u)^T (x - u) / N in float64."""
    from runia_core_b200 import _device, _ops

    rng = np.random.RandomState(3)
    x = (rng.standard_normal((n, d)) * 2 + 1).astype(np.float32)
    u = rng.standard_normal(d)
    got = _ops.shifted_covariance(_device.to_device(x), u)
    ec = EmpiricalCovariance(assume_centered=True).fit(x - u)
    np.testing.assert_allclose(got, ec.covariance_, rtol=1e-12, atol=1e-12 * np.abs(ec.covariance_).max())


@pytest.mark.parametrize("B,n", [(10, 512), (3, 130), (1, 1), (4, 7)])
def test_tril_inverse_matches_scipy(B, n):
    from scipy.linalg import solve_triangular

    from runia_core_b200 import _ops

    rng = np.random.RandomState(B * n)
    L = np.tril(rng.standard_normal((B, n, n)) / np.sqrt(n)) + np.eye(n) * (1.0 + rng.rand(B, n))[:, :, None] * np.eye(n)
    got = _ops.tril_inverse(L).cpu().numpy()
    for b in range(B):
        want = solve_triangular(L[b], np.eye(n), lower=True)
        np.testing.assert_allclose(got[b], want, rtol=1e-10, atol=1e-12 * np.abs(want).max())
        assert np.array_equal(np.triu(got[b], 1), np.zeros((n, n)))


@pytest.mark.parametrize("head_dtype,tol", [(np.float64, 2e-6), (np.float32, 1e-4)])
def test_vim_device_fit_scores_match_the_host_fit(head_dtype, tol):
    """ViM.setup with the device covariance + Jacobi eigensolver against the reference's host expressions
    (EmpiricalCovariance + np.linalg.eig, postprocessors.py:1045-1080): same alpha and scores to 1e-6 relative -- the
    residual norm depends on the span of the discarded eigenvectors only."""
    from sklearn.covariance import EmpiricalCovariance as EC

    from runia_core_b200 import inference as I

    rng = np.random.RandomState(8)
    C, d, n = 10, 128, 20_000
    mix = np.eye(d) + 0.3 * rng.standard_normal((d, d)) / np.sqrt(d)
    train = (rng.standard_normal((n, d)) @ mix * np.linspace(0.3, 2.0, d)).astype(np.float32)
    test = (1.3 * rng.standard_normal((3000, d)) @ mix).astype(np.float32)
    # a float64 head gives a float64 u and float64 host arithmetic upstream; a float32 head float32 throughout
    W = (0.1 * rng.standard_normal((C, d))).astype(head_dtype)
    b = rng.standard_normal(C).astype(head_dtype)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    p = I.ViM(flip_sign=False)
    p.setup(train, valid_feats=test, train_logits=lg(train), valid_logits=lg(test),
            final_linear_layer_params={"weight": W, "bias": b})
    got = p.postprocess(test, logits=lg(test))
    # the reference's expressions on the host
    u = -np.matmul(np.linalg.pinv(W), b)
    DIM = d // 2
    ec = EC(assume_centered=True).fit(train - u)
    ev, V = np.linalg.eig(ec.covariance_)
    NS = np.ascontiguousarray((V.T[np.argsort(ev * -1)[DIM:]]).T)
    vl_train = np.linalg.norm(np.matmul(train - u, NS), axis=-1)
    alpha = lg(train).max(axis=-1).mean() / vl_train.mean()
    from scipy.special import logsumexp

    want = -(np.linalg.norm(np.matmul(test - u, NS), axis=-1) * alpha) + logsumexp(lg(test), axis=-1)
    assert u.dtype == head_dtype
    assert abs(p.alpha - alpha) / alpha < tol
    assert np.abs(got - want).max() / np.abs(want).max() < tol
