"""GPU parity of the OoD detection metrics (-m gpu): `get_auroc_results` against the reference's golden
values (tests/unit_test_metrics.py:21-29), the oracle's restatement of torchmetrics / sklearn on seeded
inputs with ties, saturating sigmoids and scores already inside [0, 1], and at 2e7 scores through
properties (exact Mann-Whitney AUROC from an independent count, invariance under shuffling)."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from tests import refkats as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from runia_core_b200 import _ops
    from runia_core_b200.evaluation import get_auroc_results

    return _ops, get_auroc_results


def test_kat_get_auroc_results(M):
    _, get_auroc_results = M
    np.random.seed(1)
    ind = 0.5 + np.random.randn(1000)
    ood = -0.5 + np.random.randn(1000)
    res, ml = get_auroc_results("test", ind, ood, True)
    assert list(res.columns) == ["auroc", "fpr@95", "aupr", "fpr", "tpr"] and res.index[0] == "test"
    assert abs(res["fpr@95"].values[0] - K.METRICS_1D["fpr95"]) < 1e-7
    assert abs(res["aupr"].values[0] - K.METRICS_1D["aupr"]) < 1e-7
    assert abs(res["auroc"].values[0] - K.METRICS_1D["auroc"]) < 1e-7
    assert set(ml) == {"auroc", "aupr", "fpr_95"}
    fpr, tpr = np.asarray(res["fpr"].values[0]), np.asarray(res["tpr"].values[0])
    y = np.concatenate([np.ones(1000, np.int64), np.zeros(1000, np.int64)])
    rf, rt = O.binary_roc(np.concatenate([ind, ood]), y)
    assert fpr.shape == rf.shape and np.abs(fpr - rf).max() < 1e-6 and np.abs(tpr - rt).max() < 1e-6


@pytest.mark.parametrize("case", ["f64_sigmoid", "f32_sigmoid", "ties", "saturating", "unit_interval", "unbalanced"])
def test_metrics_vs_oracle(M, case):
    ops, _ = M
    rng = np.random.RandomState(sum(map(ord, case)))
    n1, n0 = 30_011, 20_003
    ind = 0.7 + rng.randn(n1)
    ood = -0.3 + 1.3 * rng.randn(n0)
    if case == "f32_sigmoid":
        ind, ood = ind.astype(np.float32), ood.astype(np.float32)
    elif case == "ties":
        ind, ood = np.round(ind * 4) / 4, np.round(ood * 4) / 4  # ~40 distinct values
    elif case == "saturating":
        ind, ood = ind * 30, ood * 30  # sigmoid(x) == 1.0 / 0.0 for many: distinct scores merge
    elif case == "unit_interval":
        ind, ood = rng.beta(5, 2, n1), rng.beta(2, 5, n0)  # already in [0, 1]: no sigmoid
    elif case == "unbalanced":
        ood = ood[:37]
    got = ops.ood_metrics(ind, ood)
    ref = O.ood_metrics(ind, ood)
    assert abs(got["auroc"] - ref[0]) < 1e-6 and abs(got["fpr95"] - ref[1]) < 1e-6 and abs(got["aupr"] - ref[2]) < 1e-6
    y = np.concatenate([np.ones(ind.size, np.int64), np.zeros(ood.size, np.int64)])
    rf, rt = O.binary_roc(np.concatenate([ind, ood]), y)
    if case == "f32_sigmoid":
        # which float32 sigmoids collide depends on the last ulp of exp (NumPy, torch and CUDA differ):
        # the number of distinct scores may move by a few in 50k; the metrics above do not
        assert abs(got["n_points"] - rf.size) <= 50
    else:
        assert got["n_points"] == rf.size
        assert np.abs(got["fpr"].cpu().numpy() - rf).max() < 1e-6 and np.abs(got["tpr"].cpu().numpy() - rt).max() < 1e-6


def test_metrics_2e7_scores(M):
    """1e7 + 1e7 float32 scores on the device: AUROC must equal the Mann-Whitney statistic counted
    independently (torch.searchsorted on the sorted OoD scores), and must not depend on the order of the
    inputs."""
    ops, _ = M
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 10_000_000
    ind = torch.sigmoid(0.5 + torch.randn(n, generator=g, device="cuda"))
    ood = torch.sigmoid(-0.5 + torch.randn(n, generator=g, device="cuda"))
    ood[: n // 10] = ind[: n // 10]  # exact cross-class ties
    got = ops.ood_metrics(ind, ood, want_curve=False)
    so = torch.sort(ood).values
    less = torch.searchsorted(so, ind, right=False).double()
    leq = torch.searchsorted(so, ind, right=True).double()
    mw = float(((less + leq) * 0.5).sum() / (float(n) * float(n)))
    assert abs(got["auroc"] - mw) < 1e-9
    perm = torch.randperm(n, generator=g, device="cuda")
    again = ops.ood_metrics(ind[perm], ood.flip(0), want_curve=False)
    assert again["auroc"] == got["auroc"] and again["fpr95"] == got["fpr95"] and abs(again["aupr"] - got["aupr"]) < 1e-12


def test_eigen_score_kat_and_config5():
    """Reference golden (unit_test_llm_uncertainty.py:69-92) through the public function, and BASELINE
    configs[4]: 10 samples x 4096-d hidden states against the oracle."""
    from runia_core_b200.llm_uncertainty import eigen_score

    np.random.seed(42)
    torch.manual_seed(42)
    hs = tuple(tuple(torch.randn(1, 10, 768) for _ in range(20)) for _ in range(5))
    got = eigen_score(hs, alpha=1e-3)
    assert isinstance(got, float) and abs(got - K.EIGEN_SCORE) < 1e-6
    rng = np.random.RandomState(42)
    E = rng.randn(10, 4096).astype(np.float32)
    E[3] = E[2]  # duplicated generation: rank drops, one more eigenvalue at alpha
    hs2 = ((None,) * 15 + (torch.from_numpy(E)[None],),)
    for alpha in (1e-3, 1e-1):
        assert abs(eigen_score(hs2, alpha=alpha) - O.eigen_score(E, alpha)) < 1e-9
    E32 = rng.randn(32, 64).astype(np.float32)
    assert abs(eigen_score(((None,) * 15 + (torch.from_numpy(E32),),), 1e-3) - O.eigen_score(E32, 1e-3)) < 1e-9


@pytest.mark.parametrize("C,n_mc", [(10, 16), (3, 5), (32, 2), (100, 8), (1000, 4)])
def test_predictive_uncertainty_vs_reference_formula(C, n_mc):
    """get_predictive_uncertainty_score against the torch expression of funcs.py:448-463 evaluated on the
    same device tensor (and the NumPy oracle), including the NaN upstream produces when a probability
    underflows to 0."""
    from runia_core_b200.inference.funcs import get_predictive_uncertainty_score

    g = torch.Generator(device="cuda").manual_seed(C * 100 + n_mc)
    n_items = 2049
    x = 3.0 * torch.randn(n_items * n_mc, C, generator=g, device="cuda")
    x[7 * n_mc, 0] = 200.0  # softmax underflow -> 0 * log 0 = NaN upstream
    ph, mi = get_predictive_uncertainty_score(x, n_mc)
    sm = torch.softmax(x, dim=1)
    st = torch.stack(torch.split(sm, n_mc))
    ep = st.mean(1)
    rph = -(ep * torch.log(ep)).sum(1)
    rmi = rph - (-(st * torch.log(st)).sum(-1)).mean(1)
    ok = torch.isfinite(rph) & torch.isfinite(rmi)
    assert ph.shape == (n_items,) and ph.dtype == torch.float32 and ph.is_cuda
    assert bool(torch.isnan(ph[7]) or torch.isnan(mi[7])) == bool(~ok[7])
    assert torch.allclose(ph[ok], rph[ok], rtol=1e-4, atol=1e-5) and torch.allclose(mi[ok], rmi[ok], rtol=1e-4, atol=2e-5)
    oh, om = O.predictive_uncertainty(x.cpu().numpy(), n_mc)
    okn = ok.cpu().numpy()
    assert np.allclose(ph.cpu().numpy()[okn], oh[okn], rtol=1e-4, atol=1e-5)
    assert np.allclose(mi.cpu().numpy()[okn], om[okn], rtol=1e-4, atol=2e-5)
    ph_c, mi_c = get_predictive_uncertainty_score(x.cpu(), n_mc)
    assert not ph_c.is_cuda and torch.equal(ph_c[okn], ph.cpu()[okn])


@pytest.mark.parametrize("shape", [(64, 512, 7, 7), (3, 5, 1, 1), (16, 64, 32, 48), (2, 1, 9, 40)])
def test_spatial_mean_vs_torch(shape):
    """get_mean_or_fullmean_ls_sample against the torch expression of utils.py:82-92 (mean over W, then H)."""
    from runia_core_b200.feature_extraction import get_mean_or_fullmean_ls_sample

    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g, device="cuda") + 0.5
    for method in ("fullmean", "mean"):
        ref = torch.mean(x, dim=3, keepdim=True)
        if method == "fullmean":
            ref = torch.mean(ref, dim=2, keepdim=True)
        ref = torch.squeeze(ref)
        got = get_mean_or_fullmean_ls_sample(x, method)
        assert got.shape == ref.shape and got.is_cuda
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
    assert not get_mean_or_fullmean_ls_sample(x.cpu(), "fullmean").is_cuda


def test_react_threshold_percentile_bit_exact():
    """The ReAct clip threshold: np.percentile(train.flatten(), p) from the device sort must be the very
    float NumPy returns (postprocessors.py:1433), at the configs[1] size (50k x 512 = 25.6M activations)."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(3)
    for n, qs in ((70_001, (0, 33.3, 90, 99.99, 100)), (25_600_000, (90, 85))):
        x = np.maximum(rng.standard_normal(n).astype(np.float32), 0)
        srt = _ops.sort_f32(x).cpu().numpy()
        if n < 1_000_000:
            assert np.array_equal(srt, np.sort(x))
        else:
            assert bool((srt[1:] >= srt[:-1]).all()) and srt[0] == x.min() and srt[-1] == x.max()
        for q in qs:
            got, ref = _ops.percentile_f32(x, q), np.percentile(x, q)
            assert got == ref, (n, q, got, ref)
    neg = np.array([3.0, -0.0, 0.0, -2.5, np.inf, -np.inf, 1e-40, -1e-40] * 9000, np.float32)
    assert np.array_equal(_ops.sort_f32(neg).cpu().numpy(), np.sort(neg))
