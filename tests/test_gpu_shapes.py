"""GPU parity (-m gpu) for the shapes the reference accepts beyond the CIFAR-10 defaults: heads with any class
count (ImageNet C = 1000, COCO C = 80) for ReAct / DICE / DICE+ReAct / ASH, GEN top-M for any C, `get_dl_h_z` for
mcd_samples_nro > 32, kNN with k > 240, any number of bank rows tying at the k-th distance, un-normalised
`FlatL2Index` vectors, Mahalanobis with hundreds of classes -- each against the oracle on seeded inputs.
Tolerance 1e-4 relative (north star) unless a tighter one is written; kNN indices and distances bit-exact."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def R():
    import runia_core_b200 as pkg

    return pkg


def _head(rng, n, d, C):
    x = np.maximum(rng.randn(n, d), 0).astype(np.float32)
    x[3] = 0.0
    x[7] *= 20.0
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    return x, W, b


@pytest.mark.parametrize("n,d,C", [(3000, 768, 1000), (2500, 1024, 80), (700, 512, 33), (20000, 2048, 257)])
def test_wide_head_tensor_path_vs_oracle(R, n, d, C):
    """postprocessors.py:1444-1474 / 1325-1354 / 1591-1621 with an ImageNet- / COCO-sized head: 256-column panels
    on the tcgen05 mainloop, online log-sum-exp across panels, ragged last panel and last row tile."""
    from runia_core_b200 import _lib, _ops

    rng = np.random.RandomState(d + C)
    x, W, b = _head(rng, n, d, C)
    Wd, bd = torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda()
    thr = float(np.percentile(x, 90))
    planes = _ops.linear_planes(Wd)
    assert planes[0].shape == (C, d)
    before = _lib.launch_count()
    got = _ops.clip_linear_lse(x, Wd, bd, clip=thr, planes=planes).cpu().numpy()
    assert _lib.launch_count() - before == 1
    ref = O.react_score(x, W, b, thr)
    assert rel_err(got, ref) < RTOL
    np.testing.assert_allclose(got, ref, rtol=3e-5, atol=3e-5)
    got = _ops.clip_linear_lse(x[:777], Wd, bd).cpu().numpy()  # no clip, planes built inside
    assert rel_err(got, O.react_score(x[:777], W, b, np.inf)) < RTOL


@pytest.mark.parametrize("n,d,C", [(600, 770, 100), (300, 8192, 10), (257, 4100, 70), (100, 30, 300)])
def test_general_head_simt_path_vs_oracle(R, n, d, C):
    """Heads the tensor-core kernels cannot take (d % 4 != 0, d > 4096) and that do not fit shared memory."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(d * 7 + C)
    x, W, b = _head(rng, n, d, C)
    Wd, bd = torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda()
    thr = float(np.percentile(x, 90))
    got = _ops.clip_linear_lse(x, Wd, bd, clip=thr).cpu().numpy()
    assert rel_err(got, O.react_score(x, W, b, thr)) < RTOL


def test_react_dice_ash_postprocessors_imagenet_head(R):
    """The registered classes end to end at C = 1000, d = 768 (no RUNIA_E_UNSUPPORTED for shapes the reference
    accepts): setup on train features, thresholds, scores vs the oracle's restatement of each class."""
    rng = np.random.RandomState(77)
    C, d, ntr, nte = 1000, 768, 4000, 1500
    means = rng.randn(C, d).astype(np.float32)
    ytr = rng.randint(0, C, ntr)
    train = np.maximum(means[ytr] + rng.randn(ntr, d), 0).astype(np.float32)
    valid = np.maximum(means[rng.randint(0, C, 500)] + rng.randn(500, d), 0).astype(np.float32)
    test = np.maximum(1.5 * rng.randn(nte, d), 0).astype(np.float32)
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    fc = {"weight": W, "bias": b}
    I = R.inference
    react = I.ReAct(flip_sign=False, react_percentile=90)
    react.setup(train, valid_feats=valid, final_linear_layer_params=fc)
    thr = O.react_threshold(train, 90)
    assert np.float32(react.activation_threshold) == np.float32(thr)
    assert rel_err(react.postprocess(test), O.react_score(test, W, b, thr)) < RTOL
    dice = I.DICE(flip_sign=False, dice_percentile=90, num_classes=C)
    dice.setup(train, valid_feats=valid, final_linear_layer_params=fc)
    mw, _ = O.dice_masked_weight(train, W, 90)
    ref = O.logsumexp(test @ mw.T + b, axis=1)
    assert rel_err(dice.postprocess(test), ref) < RTOL
    dr = I.DICEReAct(flip_sign=True, dice_percentile=90, react_percentile=90, num_classes=C)
    dr.setup(train, valid_feats=valid, final_linear_layer_params=fc)
    ref = -O.logsumexp(test.clip(max=thr) @ mw.T + b, axis=1)
    assert rel_err(dr.postprocess(test), ref) < RTOL
    ash = I.ASH(flip_sign=False, ash_percentile=85)
    ash.setup(train, valid_feats=valid, final_linear_layer_params=fc)
    ref = O.ash_score_intended(test, W, b, 85)
    assert rel_err(ash.postprocess(test), ref) < RTOL
    assert abs(ash.threshold - O.method_threshold(O.ash_score_intended(train, W, b, 85))) < 1e-3
    # RouteDICE.forward (funcs.py:182-190): the logits themselves
    lg = dice.dice_layer(test[:64]).cpu().numpy()
    np.testing.assert_allclose(lg, test[:64] @ mw.T + b, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("d,C,pct", [(768, 1000, 85), (1024, 80, 90), (2048, 10, 65), (1030, 40, 85), (130, 200, 50),
                                     (512, 10, 90), (1000, 16, 35), (64, 3, 99)])
def test_ash_any_head_vs_intended_rule(R, d, C, pct):
    """ASH-S for every head: prune kernel (k-th largest by interpolation / bisection on integer keys, ties by lowest
    index) + general head, and the fused small-head kernel.  Rows with many exact zeros, all-equal rows, quantised
    rows (ties around the cut), negative activations, 60 binades of dynamic range, two-valued and sorted rows."""
    from runia_core_b200 import _ops
    from runia_core_b200.inference.funcs import ash_s_linear_layer

    rng = np.random.RandomState(d + C + pct)
    n = 1203
    x = np.maximum(rng.randn(n, d), 0).astype(np.float32)
    x[::7] *= (rng.rand(d) < 0.08)
    x[5] = 1.0
    x[9::11] = np.round(x[9::11], 1)
    x[10::11] = np.round(x[10::11] * 2) / 2
    x[11] = -x[11] - 3.0
    x[12::97] = np.exp(9.0 * rng.randn(*x[12::97].shape)).astype(np.float32)       # values over ~60 binades
    x[13::97] = np.where(rng.rand(*x[13::97].shape) < 0.5, 0.25, 4.0)              # two values only
    x[14::97] = np.sort(x[14::97], axis=1)                                         # ascending
    x[15::97] = -np.sort(-x[15::97], axis=1)                                       # descending
    x[16] = np.arange(d, dtype=np.float32) * 1e-3                                  # equidistant
    x[17] = 1e30
    x[17, ::3] = 1e-30
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    Wd, bd = torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda()
    k_keep = d - int(np.round(d * pct / 100.0))
    shaped = ash_s_linear_layer(x, pct)
    ref_shaped = O.ash_s_intended(x, pct)
    ok = np.isfinite(ref_shaped).all(1) & np.isfinite(shaped).all(1)
    assert ok.sum() > n // 2
    assert ((shaped[ok] != 0) == (ref_shaped[ok] != 0)).all()  # the kept POSITIONS are exactly the rule's
    np.testing.assert_allclose(shaped[ok], ref_shaped[ok], rtol=2e-5, atol=1e-6)
    got = _ops.ash_linear_lse(x, Wd, bd, k_keep).cpu().numpy()
    ref = O.ash_score_intended(x, W, b, pct)
    ok = np.isfinite(ref)
    assert rel_err(got[ok], ref[ok]) < RTOL


@pytest.mark.parametrize("C,M", [(1000, 100), (1000, 1000), (200, 10), (65, 64), (5000, 7), (7000, 100)])
def test_gen_top_m_any_class_count(R, C, M):
    """GEN (postprocessors.py:650-691, funcs.py:347-375) for C > 64 with M < C; Energy / MSP from the same pass."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(C + M)
    n = 777
    lg = (3.0 * rng.randn(n, C)).astype(np.float32)
    lg[5] = 0.25                       # all classes tied
    lg[6, : C // 2] = lg[6, 0]         # half the row tied at one value
    lg[7] = np.round(lg[7])            # ties around the M-th largest
    e, m, g, _ = _ops.logit_scores(lg, gamma=0.1, M=M)
    assert rel_err(e.cpu().numpy(), O.energy_score(lg)) < RTOL
    assert rel_err(m.cpu().numpy(), O.msp_score(lg)) < RTOL
    assert rel_err(g.cpu().numpy(), O.gen_score(lg.astype(np.float64), 0.1, M)) < RTOL
    gen = R.inference.GEN(flip_sign=True, gamma=0.1, num_classes=M)
    gen.setup(lg[:100])
    assert rel_err(gen.postprocess(lg), -O.gen_score(lg.astype(np.float64), 0.1, M)) < RTOL


def test_generalized_entropy_takes_probabilities_as_given(R):
    """funcs.py:371-375 uses `probs` as they are: rows that do not sum to one (truncated / un-normalised) must not
    be re-normalised."""
    from runia_core_b200.inference.funcs import generalized_entropy

    rng = np.random.RandomState(3)
    p = rng.rand(500, 90).astype(np.float32) * 0.2      # rows sum to ~9, not 1
    p[3] = 0.0
    p[4, :10] = 1.0
    for M in (5, 90, 200):
        got = generalized_entropy(p, 0.1, M)
        assert got.dtype == np.float32
        assert rel_err(got, O.generalized_entropy(p.astype(np.float64), 0.1, M)) < RTOL
    p64 = p.astype(np.float64) / p.sum(1, keepdims=True).clip(1e-9)
    got = generalized_entropy(p64, 0.5, 10)
    assert got.dtype == np.float64 and rel_err(got, O.generalized_entropy(p64, 0.5, 10)) < RTOL


@pytest.mark.parametrize("n_mc,D", [(40, 64), (64, 100), (33, 36), (128, 32)])
def test_entropy_more_than_32_samples(R, n_mc, D):
    """get_dl_h_z(..., mcd_samples_nro > 32) (evaluation/entropy.py:41-66 accepts any count): generic kernels."""
    rng = np.random.RandomState(n_mc + D)
    n_items = 23
    z = (rng.randn(n_items, 1, D) + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32)
    z[rng.rand(n_items, n_mc, D) < 0.3] = 0.0
    z[7] = 1.25
    z = z.reshape(-1, D)
    hm, hz = R.evaluation.get_dl_h_z(z, n_mc)
    rm, rz = O.get_dl_h_z(z, n_mc, chunk=8)
    assert hz.shape == (n_items, D) and hm.shape == (n_items, 1)
    assert rel_err(hz, rz) < RTOL and rel_err(hm, rm) < RTOL


@pytest.mark.parametrize("k", [241, 500, 1016])
def test_knn_large_k_bit_exact(R, k):
    from runia_core_b200 import _ops

    rng = np.random.RandomState(k)
    bank = rng.randn(9000, 64).astype(np.float32)
    q = (rng.randn(70, 64) + 0.2 * bank[:70]).astype(np.float32)
    bn, qn = _ops.normalize_rows(bank), _ops.normalize_rows(q)
    D, I = O.flat_l2_search_tree(bn.cpu().numpy(), qn.cpu().numpy(), k)
    for planes in (True, False):  # tensor-core and FP32-SIMT candidate passes
        res = _ops.knn_search(qn, _ops.knn_bank(bn, planes=planes), k)
        assert np.array_equal(res["idx"].cpu().numpy(), I)
        assert np.array_equal(res["dist"].cpu().numpy(), D)


def test_knn_many_rows_tie_at_kth_distance(R):
    """All-zero ReLU / dropout rows normalise to identical vectors: 10,000 identical bank rows tie at every rank.
    The reference (faiss) just answers; the exhaustive pass must too (no tie-buffer limit), lowest indices first."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(9)
    d, k = 48, 50
    bank = rng.randn(14000, d).astype(np.float32)
    dup = rng.permutation(14000)[:10000]
    bank[dup] = bank[dup[0]]
    q = rng.randn(40, d).astype(np.float32)
    q[:8] = bank[dup[0]]                       # distance 0 to 10,000 rows
    q[8:16] = bank[dup[0]] + 1e-3 * rng.randn(8, d).astype(np.float32)
    bn, qn = _ops.normalize_rows(bank), _ops.normalize_rows(q)
    res = _ops.knn_search(qn, _ops.knn_bank(bn), k)
    D, I = O.flat_l2_search_tree(bn.cpu().numpy(), qn.cpu().numpy(), k)
    assert np.array_equal(res["idx"].cpu().numpy(), I)
    assert np.array_equal(res["dist"].cpu().numpy(), D)
    assert res["exhaustive_rows"] >= 8
    knn = R.inference.KNNLatentSpace()
    knn.setup(np.zeros((6000, d), np.float32))  # every training latent dead: all rows normalise to 0
    s = knn.postprocess(q)
    ref, _ = O.knn_score(q, O.normalize_rows_exact(np.zeros((6000, d), np.float32)), 50)
    assert np.array_equal(s, ref)


@pytest.mark.parametrize("cluster,spread,expect_exhaustive", [(90, 2e-3, False), (230, 1e-3, False), (700, 5e-4, True)])
def test_knn_dense_neighbourhoods_within_the_single_product_bound(R, cluster, spread, expect_exhaustive):
    """The candidate pass is one TF32 product: its distances are within ~3e-3 of the exact ones for unit-norm rows, and
    the re-rank evaluates exactly every candidate within twice that of the k-th approximate distance.  Clusters of
    DISTINCT bank rows packed far more tightly than the bound around each query: 90 rows (a prefix of the 128 kept
    survivors), 230 rows (all survivors in the band: the second pass over the lists collects it), 700 rows (denser than
    the scratch: the exhaustive pass).  Neighbours and distances must equal the float64 brute force bit for bit."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(cluster)
    d, k = 256, 50
    nq = min(300, 60_000 // cluster)  # every query has its own cluster (a query WITHOUT one would see some other
    bank = rng.randn(60_000, d).astype(np.float32)  # query's whole cluster at one distance: a band of its own)
    q = rng.randn(nq, d).astype(np.float32)
    for r in range(nq):  # query r sits next to bank rows [r * cluster, (r + 1) * cluster)
        bank[r * cluster:(r + 1) * cluster] = q[r] + (spread * np.sqrt(d) * rng.randn(cluster, d)).astype(np.float32)
    bn, qn = _ops.normalize_rows(bank), _ops.normalize_rows(q)
    res = _ops.knn_search(qn, _ops.knn_bank(bn), k, products=1)  # the single-product filter, whatever the probe says
    D, I = O.flat_l2_search_tree(bn.cpu().numpy(), qn.cpu().numpy(), k)
    assert np.array_equal(res["idx"].cpu().numpy(), I)
    assert np.array_equal(res["dist"].cpu().numpy(), D)
    assert (res["exhaustive_rows"] > 0) == expect_exhaustive, res["exhaustive_rows"]


def test_knn_dense_bank_selects_the_three_product_filter(R):
    """A bank whose rows have hundreds of neighbours inside the single-product rounding band (low intrinsic
    dimension: points on a circle embedded in 64-d) must be searched with the 3xTF32 filter -- chosen by the
    one-off density probe -- and a spread-out bank with the single product; both exact, few exhaustive rows."""
    from runia_core_b200 import _ops

    rng = np.random.RandomState(77)
    d, k = 64, 50
    t = rng.rand(40_000) * 2 * np.pi
    basis = np.linalg.qr(rng.randn(d, 2))[0]
    dense = (np.stack([np.cos(t), np.sin(t)], 1) @ basis.T).astype(np.float32)  # neighbour spacing ~1e-8 in d^2
    bank = _ops.knn_bank(_ops.normalize_rows(dense))
    assert _ops.knn_filter_products(bank, k) == 3
    tq = rng.rand(500) * 2 * np.pi
    q = _ops.normalize_rows((np.stack([np.cos(tq), np.sin(tq)], 1) @ basis.T).astype(np.float32))
    res = _ops.knn_search(q, bank, k)
    D, I = O.flat_l2_search_tree(bank.bank.cpu().numpy(), q.cpu().numpy(), k)
    assert np.array_equal(res["idx"].cpu().numpy(), I)
    assert np.array_equal(res["dist"].cpu().numpy(), D)
    spread = _ops.knn_bank(_ops.normalize_rows(rng.randn(40_000, d).astype(np.float32)))
    assert _ops.knn_filter_products(spread, k) == 1
    res1 = _ops.knn_search(q, spread, k, products=1)
    res3 = _ops.knn_search(q, spread, k, products=3)
    assert torch.equal(res1["idx"], res3["idx"]) and torch.equal(res1["dist"], res3["dist"])
    assert res1["exhaustive_rows"] == 0


def test_flat_l2_index_unnormalised_vectors(R):
    """FlatL2Index is public (stand-in for faiss.IndexFlatL2): vectors with norms ~30 and squared distances ~1e3.
    The certification bound scales with the norms, so near-ties at rank k still reach the exhaustive pass."""
    rng = np.random.RandomState(21)
    d, k = 128, 20
    bank = (3.0 * rng.randn(20000, d)).astype(np.float32)
    q = (3.0 * rng.randn(150, d)).astype(np.float32)
    # near-duplicates at distance differences far below the FP32 rounding of |q|^2 + |b|^2 - 2 q.b (~1e-3 here)
    for r in range(40):
        base = bank[100 + r]
        for t in range(30):
            bank[1000 + 30 * r + t] = base + (1e-4 * rng.randn(d)).astype(np.float32)
        q[r] = base + (0.5 * rng.randn(d)).astype(np.float32)
    index = R.inference.postprocessors.FlatL2Index(d)
    index.add(bank[:12000])
    index.add(bank[12000:])
    Dg, Ig = index.search(q, k)
    D, I = O.flat_l2_search_tree(bank, q, k)
    assert np.array_equal(Ig, I)
    assert np.array_equal(Dg, D)


def test_mahalanobis_hundreds_of_classes(R):
    """Mahalanobis / cMD with more classes than one launch holds (C = 300 > 256: class chunks, running max)."""
    rng = np.random.RandomState(31)
    C, d, n = 300, 64, 9000
    means = (2.0 * rng.randn(C, d)).astype(np.float32)
    y = rng.randint(0, C, n)
    y[y == 17] = 18  # an empty class
    x = (means[y] + rng.randn(n, d)).astype(np.float32)
    te = (means[rng.randint(0, C, 700)] + 1.3 * rng.randn(700, d)).astype(np.float32)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ma = R.inference.Mahalanobis(flip_sign=False, num_classes=C)
        ma.setup(x, train_labels=y, valid_feats=te[:100])
        cm, prec = O.mahalanobis_fit(x, y, C)
        ref = O.mahalanobis_score(te, cm, prec, C)
    assert rel_err(ma.postprocess(te), ref) < RTOL


def test_auroc_fpr95_identity_at_1e6_scores(R):
    """North star: AUROC / FPR@95 identical to 1e-6.  LaREM scores are accumulated in float32 on the GPU and in
    float64 by the reference; 1e6 + 1e6 scores make float32 ties at the 7th digit plausible, so the metric is
    compared end to end: GPU scores -> get_auroc_results vs oracle float64 scores -> oracle metrics."""
    rng = np.random.RandomState(41)
    d, n = 64, 1_000_000
    train = (0.5 + rng.randn(50_000, d)).astype(np.float32)
    ind = (0.5 + rng.randn(n, d)).astype(np.float32)
    ood = (0.2 + 1.1 * rng.randn(n, d)).astype(np.float32)
    md = R.inference.MDLatentSpace()
    md.setup(train)
    s_ind, s_ood = md.postprocess(ind), md.postprocess(ood)
    r_ind = O.md_score(ind, md.feats_mean, md.precision)
    r_ood = O.md_score(ood, md.feats_mean, md.precision)
    assert rel_err(s_ind, r_ind) < 1e-5 and rel_err(s_ood, r_ood) < 1e-5
    _, ml = R.evaluation.get_auroc_results("larem", s_ind, s_ood, return_results_for_mlflow=True)
    auroc, fpr95, aupr = O.ood_metrics(r_ind, r_ood)  # float64 reference scores, NumPy restatement of torchmetrics
    assert abs(ml["auroc"] - auroc) < 1e-6, (ml["auroc"], auroc)
    assert abs(ml["fpr_95"] - fpr95) < 1e-6, (ml["fpr_95"], fpr95)
    assert abs(ml["aupr"] - aupr) < 1e-6, (ml["aupr"], aupr)


def _ood_ctor(I, C, k, flip, gen_M=None):
    return {
        "energy": lambda: I.Energy(flip_sign=flip),
        "msp": lambda: I.MSP(flip_sign=flip),
        "gen": lambda: I.GEN(flip_sign=flip, gamma=0.1, num_classes=gen_M or C),
        "ddu": lambda: I.DDU(flip_sign=flip, num_classes=C),
        "knn": lambda: I.KNN(flip_sign=flip, k_neighbors=k),
        "mahalanobis": lambda: I.Mahalanobis(flip_sign=flip, num_classes=C),
        "vim": lambda: I.ViM(flip_sign=flip),
        "ash": lambda: I.ASH(flip_sign=flip, ash_percentile=85),
        "react": lambda: I.ReAct(flip_sign=flip, react_percentile=90),
        "dice": lambda: I.DICE(flip_sign=flip, dice_percentile=90, num_classes=C),
        "dice_react": lambda: I.DICEReAct(flip_sign=flip, dice_percentile=90, react_percentile=90, num_classes=C),
    }


def _run_fixture(R, f, names, flip, gen_M=None):
    import warnings

    C, k = int(f["num_classes"]), int(f["k"])
    fc = {"weight": f["W"], "bias": f["b"]}
    mk = _ood_ctor(R.inference, C, k, flip, gen_M)
    for name in names:
        p = mk[name]()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if name in ("energy", "msp", "gen"):
                p.setup(f["train_logits"])
                o = p.postprocess(f["ood_logits"])
            else:
                p.setup(f["train"], valid_feats=f["valid"], train_labels=f["train_labels"], train_logits=f["train_logits"],
                        valid_logits=f["valid_logits"], final_linear_layer_params=fc)
                o = p.postprocess(f["ood"], logits=f["ood_logits"])
        ref = f[f"{name}_ood"]
        assert o.dtype == ref.dtype, (name, o.dtype, ref.dtype)
        assert rel_err(o, ref) < RTOL, (name, rel_err(o, ref))
        thr = float(f[f"{name}_threshold"])
        assert abs(p.threshold - thr) < 1e-4 * max(1.0, abs(thr)), (name, p.threshold, thr)


def test_fixture_flip_sign_all_postprocessors(R, golden):
    """tests/golden/baselines_flip.npz (unmodified reference, flip_sign=True everywhere): scores AND thresholds,
    including KNN.setup's double flip, ViM.postprocess's missing flip and ASH's train-feature threshold."""
    _run_fixture(R, golden("baselines_flip"),
                 ("energy", "msp", "gen", "ddu", "knn", "mahalanobis", "vim", "ash", "react", "dice", "dice_react"), True)


def test_fixture_wide_shapes(R, golden):
    """tests/golden/wide_shapes.npz (unmodified reference): 100-class head, GEN top-10 of 100, k = 300, 40 samples."""
    w = golden("wide_shapes")
    _run_fixture(R, w, ("energy", "msp", "gen", "knn", "mahalanobis", "ash", "react", "dice", "dice_react"), False,
                 gen_M=int(w["gen_M"]))
    hm, hz = R.evaluation.get_dl_h_z(w["n40_z"], int(w["n40_n_mc"]))
    assert rel_err(hz, w["n40_h_z"]) < RTOL and rel_err(hm, w["n40_h_mvn"]) < RTOL


def test_entropy_very_wide_rows(R):
    """Row widths far beyond a latent vector (a whole flattened map handed to get_dl_h_z): no size assumption."""
    rng = np.random.RandomState(5)
    n_items, n_mc, D = 3, 16, 70_000
    z = (rng.randn(n_items, 1, D) + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32).reshape(-1, D)
    hm, hz = R.evaluation.get_dl_h_z(z, n_mc)
    rm, rz = O.get_dl_h_z(z[: 16], n_mc, chunk=1)
    assert rel_err(hz[:1], rz) < RTOL and rel_err(hm[:1], rm) < RTOL
    md = R.inference.MDLatentSpace()
    md.setup(rng.randn(100, 8).astype(np.float32))
    with pytest.raises(ValueError, match="width 8"):
        md.postprocess(rng.randn(4, 9).astype(np.float32))


def test_calculate_all_baselines_matches_reference_driver(R, golden):
    """tests/golden/all_baselines.npz: the REFERENCE's `calculate_all_baselines` (evaluation/baselines.py:713-854, run
    unmodified on its own postprocessors) over all twelve baselines, two OoD sets and an 11-column head (background
    column dropped for the labels); the product's driver on the same dictionaries must return the same arrays under
    the same keys."""
    from runia_core_b200.evaluation import calculate_all_baselines

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    f = golden("all_baselines")
    ind = {k[4:]: f[k] for k in f.files if k.startswith("in::") and not k[4:].startswith("ood_")}
    ood = {k[4:]: f[k] for k in f.files if k.startswith("in::") and k[4:].startswith("ood_")}
    names = ["vim", "msp", "raw", "knn", "energy", "ash", "gen", "react", "dice", "dice_react", "mdist", "ddu"]
    cfg = Cfg(ood_datasets=["ood_a", "ood_b"], k_neighbors=10, ash_percentile=85, gen_gamma=0.1, react_percentile=90,
              dice_percentile=90)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ind2, ood2, scores = calculate_all_baselines(baselines_names=names, ind_data_dict=ind, ood_data_dict=ood,
                                                     fc_params={"weight": f["W"], "bias": f["b"]}, cfg=cfg,
                                                     num_classes=int(f["num_classes"]))
    assert sorted(scores) == sorted(k[5:] for k in f.files if k.startswith("ood::") and not k.endswith("labels"))
    for k, v in scores.items():
        ref = f[f"ood::{k}"]
        assert v.shape == ref.shape and v.dtype == ref.dtype, (k, v.dtype, ref.dtype)
        assert rel_err(v, ref) < RTOL, (k, rel_err(v, ref))
    for k in names:
        assert rel_err(ind2[k], f[f"ind::{k}"]) < RTOL, k
    assert "train logits" not in ind2 and np.array_equal(ind2["train labels"], f["ind::train labels"])
    assert np.array_equal(ind2["valid labels"], f["ind::valid labels"])
    assert np.array_equal(ood2["ood_a labels"], f["ood::ood_a labels"])
    with pytest.raises(ValueError, match="num_classes greater than 21"):
        calculate_all_baselines(["gen"], {}, {}, None, cfg, 22)
