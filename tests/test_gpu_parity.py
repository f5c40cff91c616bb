"""GPU parity (-m gpu): the CUDA path, called through the reference-facing classes (which call the
C ABI), against (1) the reference's hard-coded golden vectors, (2) fixtures produced by the
unmodified reference source, (3) the oracle on seeded inputs.  Tolerances: 1e-4 relative
(north star) unless a tighter one is written; kNN indices bit-exact."""
import warnings

import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from tests import refkats as K
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def R():
    import runia_core_b200 as pkg

    return pkg


# ------------------------------- reference KATs through the product API --------------------
def test_kat_md_kde_10x32(R):
    tr, _, _ = K.generate_test_data(seed=42)
    te, _, _ = K.generate_test_data(seed=43)
    md = R.inference.MDLatentSpace()
    md.setup(tr)
    assert md.feats_mean.shape == (1, 32) and md.precision.shape == (32, 32) and md.centered_data.shape == tr.shape
    s = md.postprocess(te)
    assert s.dtype == np.float64 and rel_err(s, K.MD_10x32) < 1e-5  # element-wise (upstream asserts a signed sum)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        md.setup(tr)
        assert len(w) == 1 and "already trained" in str(w[0].message)
    kde = R.inference.KDELatentSpace()
    kde.setup(tr)
    s = kde.postprocess(te)
    assert s.dtype == np.float64 and rel_err(s, K.KDE_10x32) < 1e-5


def test_kat_cmd_mahalanobis_knn_10x32(R):
    tr, ytr, _ = K.generate_test_data(seed=42)
    te, _, _ = K.generate_test_data(seed=43)
    va, _, _ = K.generate_test_data(seed=44)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cmd = R.inference.cMDLatentSpace()
        cmd.setup(tr, ind_train_labels=ytr)
        s = cmd.postprocess(te, pred_labels=np.zeros(10, int))
        assert s.dtype == np.float32 and rel_err(s, K.CMD_10x32) < RTOL
        ma = R.inference.Mahalanobis(flip_sign=True, num_classes=10)
        ma.setup(tr, train_labels=ytr, valid_feats=va)
        s = ma.postprocess(te)
        assert s.dtype == np.float64 and rel_err(s, K.MAHALANOBIS_10x32) < RTOL
    knn = R.inference.KNNLatentSpace()
    knn.setup(tr)
    s = knn.postprocess(te)
    assert knn.K == 50 and s.dtype == np.float32
    assert np.array_equal(s.astype(np.float64), K.KNN_10x32)  # k > bank -> -FLT_MAX


def test_kat_energy_gen(R):
    _, _, ltr = K.generate_test_data(seed=42)
    _, _, lte = K.generate_test_data(seed=43)
    e = R.inference.Energy(flip_sign=True)
    with pytest.raises(AssertionError, match="setup\\(\\) must be called"):
        e.postprocess(lte)
    e.setup(ltr)
    assert rel_err(e.postprocess(lte), K.ENERGY_10x10) < 1e-6
    assert rel_err(e.postprocess(torch.from_numpy(lte)), K.ENERGY_10x10) < 1e-6
    g = R.inference.GEN(flip_sign=True, gamma=0.1, num_classes=10)
    g.setup(ltr)
    assert rel_err(g.postprocess(lte), K.GEN_10x10) < 1e-5


def test_kat_larem_lared_200x20(R):
    np.random.seed(1)
    x = np.random.rand(200, 20)
    md = R.inference.LaREMPostprocessor()
    md.setup(x)
    assert np.allclose(md.precision[0], K.LAREM_PRECISION_ROW0, atol=1e-6)
    assert rel_err(md.postprocess(x)[:20], K.LAREM_SCORES_20) < 1e-5
    kde = R.inference.LaREDPostprocessor()
    kde.setup(x)
    assert rel_err(kde.postprocess(x)[:20], K.LARED_SCORES_20) < 1e-5


def test_kat_entropy(R):
    np.random.seed(1)
    x = np.random.rand(3, 20)
    h = R.evaluation.single_image_entropy_calculation(x, 2)
    assert h.shape == (20,) and np.allclose(h, K.ENTROPY_SINGLE_3x20, atol=1e-5)
    torch.manual_seed(1)
    z = torch.rand(600, 20)
    h_mvn, h_z = R.evaluation.get_dl_h_z(z, 3, parallel_run=True)
    assert h_z.shape == (200, 20) and h_mvn.shape == (200, 1) and h_z.dtype == np.float64
    assert np.allclose(h_z[0], K.ENTROPY_DL_ROW0, atol=1e-5)
    same = np.full((3, 4), 0.7, np.float32)
    assert abs(R.evaluation.get_dl_h_z(same, 3)[1][0, 0] - K.ENTROPY_ALL_EQUAL) < 1e-5


def test_kat_pca(R):
    np.random.seed(1)
    ind = 0.5 + np.random.randn(1000, 20)
    ood = -0.5 + np.random.randn(1000, 20)
    tr, pca = R.apply_pca_ds_split(ind, 10)
    assert np.abs(np.asarray(tr[0], np.float64) - K.PCA_TRANSFORMED_ROW0).max() < 1e-7  # element-wise
    assert abs(float((pca.components_[0] + K.PCA_NEG_COMPONENT0).sum())) < 1e-7
    z = R.apply_pca_transform(ood, pca)
    assert z.dtype == np.float64 and rel_err(z[0], K.PCA_OOD_ROW0) < 1e-5


# ------------------------------- fixtures from the reference run ---------------------------
@pytest.mark.parametrize("name", ["latent_f32", "latent_f64"])
def test_fixture_latent(R, golden, name):
    g = golden(name)
    C, k = int(g["num_classes"]), int(g["k"])

    class Cfg:
        k_neighbors = k
        num_classes = C

    P = R.inference.postprocessors_dict
    for key, tol in (("MD", RTOL), ("cMD", RTOL), ("GMM", RTOL), ("KNN", RTOL)):
        if key == "KNN" and name != "latent_f32":
            continue
        p = P[key](cfg=Cfg())
        p._setup_flag = False
        p.setup(g["train"], ind_train_labels=g["train_labels"])
        if key == "MD":  # fitted state against the reference's own (float32 banks: the device fit of csrc/fit.cu)
            assert p.feats_mean.dtype == g["MD_feats_mean"].dtype and np.array_equal(p.feats_mean, g["MD_feats_mean"])
            assert np.abs(p.precision - g["MD_precision"]).max() <= 1e-9 * np.abs(g["MD_precision"]).max()
        for split in ("valid", "ood"):
            s = p.postprocess(g[split], pred_labels=g["valid_labels"])
            ref = g[f"{key}_{split}"]
            assert s.dtype == ref.dtype, (key, s.dtype, ref.dtype)
            assert rel_err(s, ref) < tol, (key, split, rel_err(s, ref))
    p = P["KDE"](cfg=None)
    p.setup(g["train"])
    assert rel_err(p.postprocess(g["valid"]), g["KDE_valid"]) < 1e-5
    # far-OoD queries: the reference (sklearn tree KDE) is itself off by up to nats there
    # (DESIGN.md "KDE parity"); the closed form is the oracle
    assert rel_err(p.postprocess(g["ood"]), O.kde_score(g["ood"], g["train"])) < 1e-5


def test_fixture_baselines(R, golden):
    b = golden("baselines")
    C, k = int(b["num_classes"]), int(b["k"])
    I = R.inference
    fc = {"weight": b["W"], "bias": b["b"]}
    mk = {
        "energy": lambda: I.Energy(flip_sign=False),
        "msp": lambda: I.MSP(flip_sign=False),
        "gen": lambda: I.GEN(flip_sign=False, gamma=0.1, num_classes=C),
        "ddu": lambda: I.DDU(flip_sign=False, num_classes=C),
        "knn": lambda: I.KNN(flip_sign=False, k_neighbors=k),
        "mahalanobis": lambda: I.Mahalanobis(flip_sign=False, num_classes=C),
        "vim": lambda: I.ViM(flip_sign=False),
        "ash": lambda: I.ASH(flip_sign=False, ash_percentile=85),
        "react": lambda: I.ReAct(flip_sign=False, react_percentile=90),
        "dice": lambda: I.DICE(flip_sign=False, dice_percentile=90, num_classes=C),
        "dice_react": lambda: I.DICEReAct(flip_sign=False, dice_percentile=90, react_percentile=90, num_classes=C),
    }
    for name, ctor in mk.items():
        p = ctor()
        if name in ("energy", "msp", "gen"):
            p.setup(b["train_logits"])
            v, o = p.postprocess(b["valid_logits"]), p.postprocess(b["ood_logits"])
        else:
            p.setup(b["train"], valid_feats=b["valid"], train_labels=b["train_labels"],
                    train_logits=b["train_logits"], valid_logits=b["valid_logits"],
                    final_linear_layer_params=fc)
            v = p.postprocess(b["valid"], logits=b["valid_logits"])
            o = p.postprocess(b["ood"], logits=b["ood_logits"])
        for got, split in ((v, "valid"), (o, "ood")):
            ref = b[f"{name}_{split}"]
            assert got.dtype == ref.dtype, (name, got.dtype, ref.dtype)
            assert rel_err(got, ref) < RTOL, (name, split, rel_err(got, ref))
        assert abs(p.threshold - float(b[f"{name}_threshold"])) < 1e-4 * max(1.0, abs(float(b[f"{name}_threshold"]))), name
    e = I.Energy(flip_sign=True)
    e.setup(b["train_logits"])
    assert rel_err(e.postprocess(b["valid_logits"]), b["energy_flipped_valid"]) < 1e-6


def test_fixture_entropy(R, golden):
    e = golden("entropy")
    for tag in ("n16", "n3", "n5", "n32", "n7"):
        hm, hz = R.evaluation.get_dl_h_z(e[f"{tag}_z"], int(e[f"{tag}_n_mc"]))
        assert hm.shape == e[f"{tag}_h_mvn"].shape and hz.shape == e[f"{tag}_h_z"].shape
        assert rel_err(hz, e[f"{tag}_h_z"]) < RTOL, tag
        assert rel_err(hm, e[f"{tag}_h_mvn"]) < RTOL, tag
    hm, hz = R.evaluation.get_dl_h_z(torch.from_numpy(e["ragged_z"]), int(e["ragged_n_mc"]))
    assert hz.shape == e["ragged_h_z"].shape
    assert rel_err(hz, e["ragged_h_z"]) < RTOL and rel_err(hm, e["ragged_h_mvn"]) < RTOL
    with pytest.raises(ValueError):
        R.evaluation.get_dl_h_z(e["ragged_z"], int(e["ragged_n_mc"]))


def test_fixture_pca(R, golden):
    from sklearn.decomposition import PCA

    p = golden("pca")
    for tag in ("f64", "f32"):
        pca = PCA(n_components=int(p[f"{tag}_d"]), whiten=True)
        pca.mean_, pca.components_ = p[f"{tag}_mean"], p[f"{tag}_components"]
        pca.explained_variance_ = p[f"{tag}_explained_variance"]
        z = R.apply_pca_transform(p[f"{tag}_test"], pca)
        assert z.dtype == p[f"{tag}_test_t"].dtype
        assert rel_err(z, p[f"{tag}_test_t"]) < 1e-5
        # same global-RNG state as when the reference fitted (oracle/gen_golden.py:pca_case):
        # the randomized SVD consumes np.random right after the two randn() draws
        np.random.seed(int(p[f"{tag}_seed"]))
        np.random.randn(*p[f"{tag}_train"].shape), np.random.randn(*p[f"{tag}_test"].shape)
        tr, est = R.apply_pca_ds_split(p[f"{tag}_train"], int(p[f"{tag}_d"]))
        assert rel_err(tr, p[f"{tag}_train_t"]) < 1e-5
        assert rel_err(est.transform(p[f"{tag}_test"]), p[f"{tag}_test_t"]) < 1e-5


# ------------------------------- oracle on seeded inputs at larger sizes --------------------
def test_md_config1_scale_vs_oracle(R):
    rng = np.random.RandomState(1)
    train = (0.5 + rng.randn(5000, 256)).astype(np.float32)
    test = np.concatenate([(0.5 + rng.randn(3000, 256)), (-0.5 + rng.randn(3000, 256))]).astype(np.float32)
    md = R.inference.MDLatentSpace()
    md.setup(train)
    got = md.postprocess(test)
    ref = O.md_score(test, md.feats_mean, md.precision)
    assert rel_err(got, ref) < 1e-5
    a1, f1 = O.auroc_fpr95(got[:3000], got[3000:])
    a2, f2 = O.auroc_fpr95(ref[:3000], ref[3000:])
    assert abs(a1 - a2) < 1e-6 and abs(f1 - f2) < 1e-6
    # device-resident input, same numbers
    got_dev = md.postprocess(torch.from_numpy(test).cuda())
    assert np.array_equal(got, got_dev)


def test_knn_indices_bit_exact_vs_oracle(R):
    from runia_core_b200 import _ops

    rng = np.random.RandomState(4)
    bank = rng.randn(6000, 96).astype(np.float32)
    bank[100:140] = bank[100]  # exact duplicates: ties at the k-th rank, broken by index
    q = (rng.randn(300, 96) + 0.1 * bank[rng.permutation(6000)[:300]]).astype(np.float32)
    q[:5] = bank[100]
    bn = _ops.normalize_rows(bank)
    assert np.array_equal(bn.cpu().numpy(), O.normalize_rows_exact(bank))
    qn = _ops.normalize_rows(q)
    assert np.array_equal(qn.cpu().numpy(), O.normalize_rows_exact(q))
    for k in (1, 10, 50, 120):
        res = _ops.knn_search(qn, _ops.knn_bank(bn), k, want_f64=True)
        D, I = O.flat_l2_search_tree(bn.cpu().numpy(), qn.cpu().numpy(), k)
        assert np.array_equal(res["idx"].cpu().numpy(), I), k
        assert np.array_equal(res["dist"].cpu().numpy(), D), k
        assert np.array_equal(res["kth"].cpu().numpy(), D[:, -1])
    assert res["exhaustive_rows"] >= 0


def test_knn_sharded_merge_matches_single(R):
    from runia_core_b200 import _ops

    rng = np.random.RandomState(5)
    bank = _ops.normalize_rows(rng.randn(5000, 64).astype(np.float32))
    qn = _ops.normalize_rows(rng.randn(200, 64).astype(np.float32))
    full = _ops.knn_search(qn, _ops.knn_bank(bank), 50)
    parts_d, parts_i = [], []
    for lo, hi in ((0, 1700), (1700, 3400), (3400, 5000)):
        r = _ops.knn_search(qn, _ops.knn_bank(bank[lo:hi].contiguous(), idx_offset=lo), 50, want_f64=True)
        parts_d.append(r["dist64"])
        parts_i.append(r["idx"])
    d, i, kth = _ops.topk_merge(torch.stack(parts_d), torch.stack(parts_i))
    assert torch.equal(i, full["idx"]) and torch.equal(d, full["dist"]) and torch.equal(kth, full["kth"])


def test_kde_vs_oracle_mid(R):
    rng = np.random.RandomState(6)
    bank = (0.5 + rng.randn(4000, 64)).astype(np.float32)
    q = np.concatenate([0.5 + rng.randn(200, 64), -0.5 + rng.randn(200, 64)]).astype(np.float32)
    kde = R.inference.KDELatentSpace()
    kde.setup(bank)
    assert rel_err(kde.postprocess(q), O.kde_score(q, bank)) < 1e-5


def test_entropy_mid_vs_oracle(R):
    rng = np.random.RandomState(7)
    n_items, n_mc, D = 300, 16, 512
    base = rng.randn(n_items, 1, D)
    z = (base + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32)
    z[rng.rand(n_items, n_mc, D) < 0.4] = 0.0
    z = z.reshape(-1, D)
    hm, hz = R.evaluation.get_dl_h_z(z, n_mc)
    rm, rz = O.get_dl_h_z(z, n_mc, chunk=16)
    assert rel_err(hz, rz) < RTOL and rel_err(hm, rm) < RTOL


def test_edge_cases(R):
    md = R.inference.MDLatentSpace()
    md.setup(np.random.RandomState(0).randn(50, 7).astype(np.float32))
    assert md.postprocess(np.zeros((0, 7), np.float32)).shape == (0,)
    assert md.postprocess(np.random.RandomState(1).randn(1, 7).astype(np.float32)).shape == (1,)
    with pytest.raises(AssertionError, match="2 dimensional"):
        md.postprocess(np.zeros(7, np.float32))
    with pytest.raises(ValueError, match="id_labels not provided"):
        R.inference.cMDLatentSpace().setup(np.zeros((4, 3), np.float32))
    with pytest.raises(AssertionError, match="valid_feats must be provided"):
        R.inference.KNN(flip_sign=False, k_neighbors=3).setup(np.zeros((4, 3), np.float32))
    with pytest.raises(ValueError, match="scores must be a dict or ndarray"):
        R.inference.Energy(flip_sign=True).flip_sign_fn([1.0])


@pytest.mark.parametrize("d,pct", [(512, 85), (512, 10), (1024, 85), (1000, 90), (130, 65)])
def test_ash_react_heads_vs_oracle(R, d, pct):
    """ASH-S (histogram select for full 512 / 1024 rows, radix select with early exit / tie path
    otherwise) and the ReAct / DICE head at widths that hit the one-chunk, two-chunk and unaligned
    paths; rows with many exact zeros put the k-th largest activation into a tie, quantised rows
    put ties into the deciding histogram bin (kept by lowest index)."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(d)
    n, C = 3001, 10
    x = np.maximum(rng.randn(n, d), 0).astype(np.float32)
    x[::7] *= (rng.rand(d) < 0.08)  # very sparse rows: fewer positives than kept elements
    x[5] = 1.0                      # all equal: ties between non-zero activations
    x[9::11] = np.round(x[9::11], 1)      # few distinct values: a handful of ties around the threshold
    x[10::11] = np.round(x[10::11] * 2) / 2  # very few distinct values: dozens of ties (exact path)
    x[11] = -x[11] - 3.0                  # negative activations
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    Wd, bd = torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda()
    k_keep = d - int(np.round(d * pct / 100.0))
    got = _ops.ash_linear_lse(x, Wd, bd, k_keep).cpu().numpy()
    # Upstream (funcs.py:249-252) scatters np.partition's values to np.argpartition's indices; the
    # two agree element for element on the reference's own fixtures (tests/golden/baselines.npz), but
    # NumPy does not promise it and on wide rows the values can land permuted among the kept
    # positions.  The check here is the intended rule: the k largest activations stay where they are
    # (ties: lowest index), everything else is zeroed, then exp(s1 / s2) scaling.
    ref = np.empty(n, np.float64)
    for r in range(n):
        keep = np.lexsort((np.arange(d), -x[r]))[:k_keep]
        kept = np.zeros(d, np.float32)
        kept[keep] = x[r, keep]
        with np.errstate(all="ignore"):
            sc = np.exp(x[r].sum(dtype=np.float32) / kept.sum(dtype=np.float32))
        ref[r] = O.logsumexp(sc * (kept @ W.T) + b)
    ok = np.isfinite(ref)
    assert ok.sum() > n // 2 and rel_err(got[ok], ref[ok]) < RTOL
    thr = float(np.percentile(x, 90))
    got = _ops.clip_linear_lse(x, Wd, bd, clip=thr).cpu().numpy()
    assert rel_err(got, O.react_score(x, W, b, thr)) < RTOL
    got = _ops.clip_linear_lse(x[:1], Wd, bd).cpu().numpy()  # odd row count, no clip
    assert rel_err(got, O.react_score(x[:1], W, b, np.inf)) < RTOL


@pytest.mark.parametrize("D", [1024, 100, 36, 30])
def test_entropy_widths_vs_oracle(R, D):
    """entropy16_kernel at the RoI width of configs[2] (1024), at widths that end inside a 64-dimension
    step (100, 36) and at a width that falls back to the warp-per-item kernel (30: D % 4 != 0)."""
    rng = np.random.RandomState(D)
    n_items, n_mc = 257, 16
    z = (rng.randn(n_items, 1, D) + 0.2 * rng.randn(n_items, n_mc, D)).astype(np.float32)
    z[rng.rand(n_items, n_mc, D) < 0.3] = 0.0
    z = z.reshape(-1, D)
    hm, hz = R.evaluation.get_dl_h_z(z, n_mc)
    rm, rz = O.get_dl_h_z(z, n_mc, chunk=32)
    assert hz.shape == (n_items, D) and hm.shape == (n_items, 1)
    assert rel_err(hz, rz) < RTOL and rel_err(hm, rm) < RTOL


@pytest.mark.parametrize("d,C,clip", [(512, 10, 1.0), (1000, 21, np.inf), (36, 32, 0.5), (2048, 10, 1.2)])
def test_react_head_tensor_path_vs_oracle(R, d, C, clip):
    """ReAct / DICE head through the narrow-panel tcgen05 kernel (>= 16384 rows, d % 4 == 0, C <= 32): ragged last
    row tile, K tails, every class count up to the panel width; same rows through the one-warp-per-row kernel."""
    from runia_core_b200 import _lib, _ops

    rng = np.random.RandomState(d + C)
    n = 16384 + 777
    x = np.maximum(rng.randn(n, d), 0).astype(np.float32)
    x[3] = 0.0
    x[7] *= 50.0
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    Wd, bd = torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda()
    before = _lib.launch_count()
    got = _ops.clip_linear_lse(x, Wd, bd, clip=clip).cpu().numpy()
    assert _lib.launch_count() - before == 2  # split of the padded planes + the tensor kernel
    ref = O.react_score(x, W, b, clip)
    assert rel_err(got, ref) < RTOL
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)
    simt = _ops.clip_linear_lse(x[:5000], Wd, bd, clip=clip).cpu().numpy()
    np.testing.assert_allclose(simt, got[:5000], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("n_items,D,n_mc", [(1300, 256, 32), (601, 132, 32), (3, 4, 32), (900, 192, 24), (700, 128, 17)])
def test_entropy_n32_many_items_vs_oracle(R, n_items, D, n_mc):
    """entropy32_kernel (four warps per item, the reference's default mcd_samples_nro = 32; 17 .. 31 samples with +inf
    sentinel rows): more items than resident CTAs, so that the tile ring runs across item boundaries and both Chebyshev
    tables are reused; ragged last tile; duplicates (min_dist clamp) and an all-equal item."""
    rng = np.random.RandomState(n_items + D)
    z = (rng.randn(n_items, 1, D) + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32)
    z[rng.rand(n_items, n_mc, D) < 0.3] = 0.0
    z[min(7, n_items - 1)] = 1.25
    z = z.reshape(-1, D)
    hm, hz = R.evaluation.get_dl_h_z(z, n_mc)
    rm, rz = O.get_dl_h_z(z, n_mc, chunk=64)
    assert hz.shape == (n_items, D) and hm.shape == (n_items, 1)
    assert rel_err(hz, rz) < RTOL and rel_err(hm, rm) < RTOL
    np.testing.assert_allclose(hz, rz, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(hm, rm, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("n_mc", [16, 32])
def test_entropy_without_joint_estimate_matches_with(R, n_mc):
    """want_joint=False (per-dimension entropies only) skips the pair maxima and the item end: h_z must not change,
    on device-resident samples and for more items than resident CTAs / warps."""
    import torch

    from runia_core_b200 import _ops

    g = torch.Generator(device="cuda").manual_seed(n_mc)
    z = torch.randn(2100 * n_mc, 192, generator=g, device="cuda")
    hm, hz = _ops.mcd_entropy(z, n_mc)
    hz = hz.clone()
    hm0, hz0 = _ops.mcd_entropy(z, n_mc, want_joint=False)
    assert hm is not None and hm0 is None
    assert torch.equal(hz, hz0)


@pytest.mark.parametrize("n_mc,D", [(32, 512), (32, 100), (20, 36), (10, 512), (8, 65), (6, 512), (16, 510),
                                    (17, 512), (31, 260), (24, 64), (25, 30)])
def test_entropy_any_n_mc_vs_oracle(R, n_mc, D):
    """entropy_np_kernel: every n_mc in [6, 32] (k = 5; 32 is the reference's default mcd_samples_nro), power-of-two
    padding with +inf sentinels, ragged last 32-dimension step, exact duplicates (min_dist clamp)."""
    rng = np.random.RandomState(n_mc * 1000 + D)
    n_items = 131
    z = (rng.randn(n_items, 1, D) + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32)
    z[rng.rand(n_items, n_mc, D) < 0.3] = 0.0
    z[7] = 1.25  # an item whose samples are all equal
    z = z.reshape(-1, D)
    hm, hz = R.evaluation.get_dl_h_z(z, n_mc)
    rm, rz = O.get_dl_h_z(z, n_mc, chunk=32)
    assert hz.shape == (n_items, D) and hm.shape == (n_items, 1)
    assert rel_err(hz, rz) < RTOL and rel_err(hm, rm) < RTOL
    np.testing.assert_allclose(hz, rz, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(hm, rm, rtol=1e-5, atol=1e-4)
