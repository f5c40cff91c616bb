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
