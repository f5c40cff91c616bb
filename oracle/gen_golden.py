"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
source (loaded by oracle/ref_loader.py from /root/reference) on seeded synthetic inputs.

Run in the build container only:   python oracle/gen_golden.py
The fixtures hold inputs AND the reference's outputs, so that the oracle (CPU, -m "not gpu")
and the CUDA path (-m gpu) are both checked against the real reference on a box where
/root/reference does not exist.

DICE / DICEReAct: the reference calls `.cuda()` unconditionally (inference/funcs.py:180,185).
To run it on this CPU-only container the generator makes `Tensor.cuda()` / `Module.cuda()` the
identity for this process; the reference source itself is untouched.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def latent_case(ref, seed, n_train, n_test, d, n_classes, dtype, k):
    rng = np.random.RandomState(seed)
    centers = rng.randn(n_classes, d) * 0.7
    ytr = rng.randint(0, n_classes, n_train)
    train = (0.5 + centers[ytr] + rng.randn(n_train, d)).astype(dtype)
    yte = rng.randint(0, n_classes, n_test)
    valid = (0.5 + centers[yte] + rng.randn(n_test, d)).astype(dtype)
    ood = (-0.5 + 1.3 * rng.randn(n_test, d)).astype(dtype)
    out = dict(train=train, train_labels=ytr, valid=valid, valid_labels=yte, ood=ood,
               num_classes=n_classes, k=k)
    cfg = ref.DictConfig(k_neighbors=k, num_classes=n_classes)
    for name in ("KDE", "MD", "cMD", "KNN", "GMM"):
        if name == "KNN" and dtype != np.float32:
            continue  # faiss only takes float32 (postprocessors.py:396-397)
        p = ref.pp.postprocessors_dict[name](cfg=cfg)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            p.setup(train, ind_train_labels=ytr)
            out[f"{name}_valid"] = np.asarray(p.postprocess(valid, pred_labels=yte))
            out[f"{name}_ood"] = np.asarray(p.postprocess(ood, pred_labels=yte))
        if name == "MD":
            out["MD_feats_mean"] = p.feats_mean
            out["MD_precision"] = p.precision
        if name == "KNN":
            out["KNN_activation_log"] = p.activation_log
    return out


def baselines_case(ref, seed, n_train, n_test, d, C, k):
    rng = np.random.RandomState(seed)
    centers = rng.randn(C, d)
    ytr = rng.randint(0, C, n_train)
    train = np.maximum(centers[ytr] + rng.randn(n_train, d), 0).astype(np.float32) + \
        (0.05 * rng.rand(n_train, d)).astype(np.float32)
    yva = rng.randint(0, C, n_test)
    valid = np.maximum(centers[yva] + rng.randn(n_test, d), 0).astype(np.float32) + \
        (0.05 * rng.rand(n_test, d)).astype(np.float32)
    ood = np.maximum(1.5 * rng.randn(n_test, d), 0).astype(np.float32) + \
        (0.05 * rng.rand(n_test, d)).astype(np.float32)
    W = (0.2 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    out = dict(train=train, train_labels=ytr, valid=valid, ood=ood, W=W, b=b,
               train_logits=lg(train), valid_logits=lg(valid), ood_logits=lg(ood),
               num_classes=C, k=k, gamma=0.1, ash_percentile=85, react_percentile=90,
               dice_percentile=90)
    fc = {"weight": W, "bias": b}
    P = ref.pp
    mk = {
        "energy": lambda: P.Energy(flip_sign=False),
        "msp": lambda: P.MSP(flip_sign=False),
        "gen": lambda: P.GEN(flip_sign=False, gamma=0.1, num_classes=C),
        "ddu": lambda: P.DDU(flip_sign=False, num_classes=C),
        "knn": lambda: P.KNN(flip_sign=False, k_neighbors=k),
        "mahalanobis": lambda: P.Mahalanobis(flip_sign=False, num_classes=C),
        "vim": lambda: P.ViM(flip_sign=False),
        "ash": lambda: P.ASH(flip_sign=False, ash_percentile=85),
        "react": lambda: P.ReAct(flip_sign=False, react_percentile=90),
        "dice": lambda: P.DICE(flip_sign=False, dice_percentile=90, num_classes=C),
        "dice_react": lambda: P.DICEReAct(flip_sign=False, dice_percentile=90,
                                           react_percentile=90, num_classes=C),
    }
    logit_methods = ("energy", "msp", "gen")
    for name, ctor in mk.items():
        p = ctor()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if name in logit_methods:
                p.setup(out["train_logits"])
                v, o = p.postprocess(out["valid_logits"]), p.postprocess(out["ood_logits"])
            else:
                p.setup(train, valid_feats=valid, train_labels=ytr,
                        train_logits=out["train_logits"], valid_logits=out["valid_logits"],
                        final_linear_layer_params=fc)
                v = p.postprocess(valid, logits=out["valid_logits"])
                o = p.postprocess(ood, logits=out["ood_logits"])
        out[f"{name}_valid"], out[f"{name}_ood"] = np.asarray(v), np.asarray(o)
        out[f"{name}_threshold"] = np.float64(p.threshold)
        if name == "vim":
            out["vim_u"], out["vim_NS"], out["vim_alpha"] = p.u, p.NS, np.float64(p.alpha)
            out["vim_DIM"] = p.DIM
        if name == "react":
            out["react_activation_threshold"] = np.float64(p.activation_threshold)
        if name == "dice":
            out["dice_masked_w"] = p.dice_layer.masked_w.cpu().numpy()
            out["dice_thresh"] = np.float64(p.dice_layer.thresh)
        if name == "mahalanobis":
            out["mahalanobis_class_mean"] = p.class_mean
            out["mahalanobis_precision"] = p.precision
    # flip_sign variant (OodPostprocessor.flip_sign_fn, abstract_classes.py:160-187)
    p = P.Energy(flip_sign=True)
    p.setup(out["train_logits"])
    out["energy_flipped_valid"] = p.postprocess(out["valid_logits"])
    out["energy_flipped_threshold"] = np.float64(p.threshold)
    return out


def flip_case(ref, seed, n_train, n_test, d, C, k):
    """Every OodPostprocessor with flip_sign=True: scores and the threshold setup() derives (KNN.setup flips the
    already flipped validation scores a second time, postprocessors.py:852-854; ViM.postprocess never flips,
    :1082-1112; ASH thresholds on the train features, :1185)."""
    rng = np.random.RandomState(seed)
    centers = rng.randn(C, d)
    ytr = rng.randint(0, C, n_train)
    mk_x = lambda y, n: (np.maximum(centers[y] + rng.randn(n, d), 0) + 0.05 * rng.rand(n, d)).astype(np.float32)  # noqa: E731
    train = mk_x(ytr, n_train)
    yva = rng.randint(0, C, n_test)
    valid = mk_x(yva, n_test)
    ood = (np.maximum(1.5 * rng.randn(n_test, d), 0) + 0.05 * rng.rand(n_test, d)).astype(np.float32)
    W = (0.2 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    out = dict(train=train, train_labels=ytr, valid=valid, ood=ood, W=W, b=b, train_logits=lg(train),
               valid_logits=lg(valid), ood_logits=lg(ood), num_classes=C, k=k)
    fc = {"weight": W, "bias": b}
    P = ref.pp
    mk = {
        "energy": lambda: P.Energy(flip_sign=True),
        "msp": lambda: P.MSP(flip_sign=True),
        "gen": lambda: P.GEN(flip_sign=True, gamma=0.1, num_classes=C),
        "ddu": lambda: P.DDU(flip_sign=True, num_classes=C),
        "knn": lambda: P.KNN(flip_sign=True, k_neighbors=k),
        "mahalanobis": lambda: P.Mahalanobis(flip_sign=True, num_classes=C),
        "vim": lambda: P.ViM(flip_sign=True),
        "ash": lambda: P.ASH(flip_sign=True, ash_percentile=85),
        "react": lambda: P.ReAct(flip_sign=True, react_percentile=90),
        "dice": lambda: P.DICE(flip_sign=True, dice_percentile=90, num_classes=C),
        "dice_react": lambda: P.DICEReAct(flip_sign=True, dice_percentile=90, react_percentile=90, num_classes=C),
    }
    for name, ctor in mk.items():
        p = ctor()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if name in ("energy", "msp", "gen"):
                p.setup(out["train_logits"])
                o = p.postprocess(out["ood_logits"])
            else:
                p.setup(train, valid_feats=valid, train_labels=ytr, train_logits=out["train_logits"],
                        valid_logits=out["valid_logits"], final_linear_layer_params=fc)
                o = p.postprocess(ood, logits=out["ood_logits"])
        out[f"{name}_ood"] = np.asarray(o)
        out[f"{name}_threshold"] = np.float64(p.threshold)
    return out


def wide_case(ref, seed):
    """Shapes beyond the CIFAR-10 defaults: a 100-class head (ReAct / DICE / DICE+ReAct / ASH / Energy / MSP /
    GEN with M < C / Mahalanobis), kNN with k = 300, get_dl_h_z with 40 MC samples."""
    rng = np.random.RandomState(seed)
    C, d, n_train, n_test = 100, 64, 1200, 96
    centers = rng.randn(C, d)
    ytr = rng.randint(0, C, n_train)
    ytr[:C] = np.arange(C)  # every class present
    mk_x = lambda y, n: (np.maximum(centers[y] + rng.randn(n, d), 0) + 0.05 * rng.rand(n, d)).astype(np.float32)  # noqa: E731
    train, valid = mk_x(ytr, n_train), mk_x(rng.randint(0, C, n_test), n_test)
    ood = (np.maximum(1.5 * rng.randn(n_test, d), 0) + 0.05 * rng.rand(n_test, d)).astype(np.float32)
    W = (0.2 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    out = dict(train=train, train_labels=ytr, valid=valid, ood=ood, W=W, b=b, train_logits=lg(train),
               valid_logits=lg(valid), ood_logits=lg(ood), num_classes=C, gen_M=10, k=300)
    fc = {"weight": W, "bias": b}
    P = ref.pp
    mk = {
        "energy": lambda: P.Energy(flip_sign=False),
        "msp": lambda: P.MSP(flip_sign=False),
        "gen": lambda: P.GEN(flip_sign=False, gamma=0.1, num_classes=10),
        "knn": lambda: P.KNN(flip_sign=False, k_neighbors=300),
        "mahalanobis": lambda: P.Mahalanobis(flip_sign=False, num_classes=C),
        "ash": lambda: P.ASH(flip_sign=False, ash_percentile=85),
        "react": lambda: P.ReAct(flip_sign=False, react_percentile=90),
        "dice": lambda: P.DICE(flip_sign=False, dice_percentile=90, num_classes=C),
        "dice_react": lambda: P.DICEReAct(flip_sign=False, dice_percentile=90, react_percentile=90, num_classes=C),
    }
    for name, ctor in mk.items():
        p = ctor()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if name in ("energy", "msp", "gen"):
                p.setup(out["train_logits"])
                o = p.postprocess(out["ood_logits"])
            else:
                p.setup(train, valid_feats=valid, train_labels=ytr, train_logits=out["train_logits"],
                        valid_logits=out["valid_logits"], final_linear_layer_params=fc)
                o = p.postprocess(ood, logits=out["ood_logits"])
        out[f"{name}_ood"] = np.asarray(o)
        out[f"{name}_threshold"] = np.float64(p.threshold)
    n_mc, n_items, D = 40, 5, 21
    z = (rng.randn(n_items, 1, D) + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32)
    z[rng.rand(n_items, n_mc, D) < 0.4] = 0.0
    z = z.reshape(n_items * n_mc, D)
    h_mvn, h_z = ref.entropy.get_dl_h_z(z, n_mc, parallel_run=False)
    out["n40_z"], out["n40_n_mc"], out["n40_h_mvn"], out["n40_h_z"] = z, n_mc, h_mvn, h_z
    return out


def all_baselines_case(ref, seed):
    """The reference's own driver, `calculate_all_baselines` (evaluation/baselines.py:713-854), over every baseline
    it knows, two OoD sets, 10 classes + a background logit column (11 columns: `get_labels_from_logits` drops the
    last one, baselines.py:645-676).  Inputs and every "<ood> <baseline>" / InD score array are stored."""
    rng = np.random.RandomState(seed)
    C, d, n_train, n_valid, n_ood = 10, 48, 1500, 120, 90
    centers = rng.randn(C, d)
    mk_x = lambda y, n: (np.maximum(centers[y] + rng.randn(n, d), 0) + 0.05 * rng.rand(n, d)).astype(np.float32)  # noqa: E731
    ytr = rng.randint(0, C, n_train)
    train, valid = mk_x(ytr, n_train), mk_x(rng.randint(0, C, n_valid), n_valid)
    oods = {"ood_a": (np.maximum(1.5 * rng.randn(n_ood, d), 0) + 0.05 * rng.rand(n_ood, d)).astype(np.float32),
            "ood_b": (np.maximum(0.3 + rng.randn(n_ood, d), 0)).astype(np.float32)}
    W = (0.2 * rng.randn(C + 1, d)).astype(np.float32)
    b = rng.randn(C + 1).astype(np.float32)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    ind = {"train features": train, "valid features": valid, "train logits": lg(train), "valid logits": lg(valid)}
    ood = {}
    for name, x in oods.items():
        ood[f"{name} features"], ood[f"{name} logits"] = x, lg(x)
    out = {f"in::{k}": v for k, v in {**ind, **ood}.items()}
    out["W"], out["b"], out["num_classes"] = W, b, C + 1
    names = ["vim", "msp", "raw", "knn", "energy", "ash", "gen", "react", "dice", "dice_react", "mdist", "ddu"]
    cfg = ref.DictConfig(ood_datasets=list(oods), k_neighbors=10, ash_percentile=85, gen_gamma=0.1, react_percentile=90,
                         dice_percentile=90)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ind2, ood2, scores = ref.baselines.calculate_all_baselines(
            baselines_names=names, ind_data_dict=dict(ind), ood_data_dict=dict(ood), fc_params={"weight": W, "bias": b},
            cfg=cfg, num_classes=C + 1)
    for k in names:
        out[f"ind::{k}"] = np.asarray(ind2[k])
    for k, v in scores.items():
        out[f"ood::{k}"] = np.asarray(v)
    out["ind::train labels"], out["ind::valid labels"] = ind2["train labels"], ind2["valid labels"]
    out["ood::ood_a labels"] = ood2["ood_a labels"]
    return out


def entropy_case(ref, seed):
    rng = np.random.RandomState(seed)
    out = {}
    for tag, n_mc, n_items, D in (("n16", 16, 12, 40), ("n3", 3, 9, 20), ("n5", 5, 6, 17),
                                  ("n32", 32, 4, 24), ("n7", 7, 5, 33)):
        base = rng.randn(n_items, 1, D)
        z = (base + 0.1 * rng.randn(n_items, n_mc, D)).astype(np.float32)
        mask = rng.rand(n_items, n_mc, D) < 0.4  # DropBlock-like zeros -> exact duplicates
        z[mask] = 0.0
        z[0] = 0.25  # all-equal item: every distance hits the min_dist clamp
        z = z.reshape(n_items * n_mc, D)
        h_mvn, h_z = ref.entropy.get_dl_h_z(z, n_mc, parallel_run=False)
        out[f"{tag}_z"], out[f"{tag}_n_mc"] = z, n_mc
        out[f"{tag}_h_mvn"], out[f"{tag}_h_z"] = h_mvn, h_z
    # Tensor path tolerates a short last chunk (entropy.py:56-58)
    z = torch.from_numpy(rng.rand(6 * 8 + 7, 10).astype(np.float32))
    h_mvn, h_z = ref.entropy.get_dl_h_z(z, 8, parallel_run=False)
    out["ragged_z"], out["ragged_n_mc"] = z.numpy(), 8
    out["ragged_h_mvn"], out["ragged_h_z"] = h_mvn, h_z
    return out


def pca_case(ref, seed):
    out = {}
    for tag, dtype, n, D0, d in (("f64", np.float64, 400, 20, 10), ("f32", np.float32, 500, 48, 16)):
        np.random.seed(seed)
        train = (0.5 + np.random.randn(n, D0)).astype(dtype)
        test = (-0.5 + np.random.randn(64, D0)).astype(dtype)
        tr, pca = ref.dimred.apply_pca_ds_split(train, d)
        te = ref.dimred.apply_pca_transform(test, pca)
        out.update({f"{tag}_train": train, f"{tag}_test": test, f"{tag}_train_t": tr,
                    f"{tag}_test_t": te, f"{tag}_components": pca.components_,
                    f"{tag}_mean": pca.mean_, f"{tag}_explained_variance": pca.explained_variance_,
                    f"{tag}_seed": seed, f"{tag}_d": d})
    # no-whiten variant
    np.random.seed(seed)
    train = 0.5 + np.random.randn(300, 12)
    tr, pca = ref.dimred.apply_pca_ds_split(train, 5, svd_solver="full", whiten=False)
    out.update(dict(nw_train=train, nw_train_t=tr, nw_components=pca.components_, nw_mean=pca.mean_))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    np.savez_compressed(os.path.join(OUT, "latent_f32.npz"),
                        **latent_case(ref, 11, 600, 96, 24, 4, np.float32, 7))
    np.savez_compressed(os.path.join(OUT, "latent_f64.npz"),
                        **latent_case(ref, 12, 400, 64, 16, 3, np.float64, 5))
    np.savez_compressed(os.path.join(OUT, "baselines.npz"),
                        **baselines_case(ref, 21, 900, 128, 32, 5, 10))
    np.savez_compressed(os.path.join(OUT, "entropy.npz"), **entropy_case(ref, 31))
    np.savez_compressed(os.path.join(OUT, "pca.npz"), **pca_case(ref, 1))
    np.savez_compressed(os.path.join(OUT, "baselines_flip.npz"), **flip_case(ref, 22, 700, 96, 32, 5, 10))
    np.savez_compressed(os.path.join(OUT, "wide_shapes.npz"), **wide_case(ref, 23))
    np.savez_compressed(os.path.join(OUT, "all_baselines.npz"), **all_baselines_case(ref, 24))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
