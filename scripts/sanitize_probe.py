"""Small invocations of every kernel family (one row tile or two each) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_probe.py
Results are compared with the plain PyTorch expression so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runia_core_b200 as R  # noqa: E402
from runia_core_b200 import _ops  # noqa: E402


def main():
    rng = np.random.RandomState(0)
    dev = torch.device("cuda", 0)
    ok = []
    # tcgen05 row scorers: LaREM, PCA, ViM, class-conditional, GMM (one ragged 256-row tile + a second tile)
    d = 64
    train = rng.randn(600, d).astype(np.float32)
    x = rng.randn(300, d).astype(np.float32)
    md = R.inference.MDLatentSpace()
    md.setup(train)
    s = md.postprocess(x)
    df = x.astype(np.float64) - md.feats_mean
    ok.append(("larem", np.allclose(s, -np.einsum("ij,jk,ik->i", df, md.precision, df), rtol=1e-4)))
    np.random.seed(1)
    tr, pca = R.apply_pca_ds_split(train, 16)
    z = R.apply_pca_transform(x, pca)
    ref = (x - pca.mean_) @ pca.components_.T / np.sqrt(pca.explained_variance_)
    ok.append(("pca", np.allclose(z, ref, rtol=1e-4, atol=1e-4)))
    y = rng.randint(0, 4, 600)
    ma = R.inference.Mahalanobis(flip_sign=False, num_classes=4)
    ma.setup(train, train_labels=y, valid_feats=x)
    ok.append(("classcond", np.isfinite(ma.postprocess(x)).all()))
    ddu = R.inference.DDU(flip_sign=False, num_classes=4)
    ddu.setup(train, train_labels=y, valid_feats=x)
    ok.append(("gmm", np.isfinite(ddu.postprocess(x)).all()))
    # kNN (seed + candidates + re-rank + exhaustive pass) and KDE
    bank = _ops.normalize_rows(rng.randn(5000, d).astype(np.float32))
    q = _ops.normalize_rows(rng.randn(300, d).astype(np.float32))
    q[:3] = bank[:3]
    res = _ops.knn_search(q, _ops.knn_bank(bank), 10)
    d2 = torch.cdist(q.double(), bank.double()).pow(2)
    ok.append(("knn", torch.equal(res["idx"], d2.topk(10, dim=1, largest=False).indices.sort(1).values) or
               bool((res["idx"].sort(1).values == d2.topk(10, dim=1, largest=False).indices.sort(1).values).all())))
    kde = R.inference.KDELatentSpace()
    kde.setup(train)
    sk = kde.postprocess(x)
    t = -0.5 * ((x[:, None, :].astype(np.float64) - train[None].astype(np.float64)) ** 2).sum(-1)
    refk = np.log(np.exp(t - t.max(1, keepdims=True)).sum(1)) + t.max(1) - np.log(600) - 0.5 * d * np.log(2 * np.pi)
    ok.append(("kde", np.allclose(sk, refk, rtol=1e-4)))
    # heads: narrow tensor path, wide tensor path, SIMT, ASH; logits; entropy; metrics
    for n, C in ((16384 + 300, 10), (300, 80)):
        xf = np.maximum(rng.randn(n, d), 0).astype(np.float32)
        W, b = (0.1 * rng.randn(C, d)).astype(np.float32), rng.randn(C).astype(np.float32)
        Wd, bd = torch.from_numpy(W).to(dev), torch.from_numpy(b).to(dev)
        got = _ops.clip_linear_lse(xf, Wd, bd, clip=1.0).cpu().numpy()
        lg = np.minimum(xf, 1.0) @ W.T + b
        ok.append((f"head_c{C}", np.allclose(got, np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1)) + lg.max(1), rtol=1e-4)))
        ok.append((f"ash_c{C}", np.isfinite(_ops.ash_linear_lse(xf[:300], Wd, bd, 10).cpu().numpy()).all()))
    lgts = rng.randn(1000, 10).astype(np.float32)
    e, m, g, _ = _ops.logit_scores(lgts, gamma=0.1, M=10)
    ok.append(("logits", np.allclose(e.cpu().numpy(), np.log(np.exp(lgts).sum(1)), rtol=1e-4)))
    zz = (rng.randn(40, 1, 64) + 0.1 * rng.randn(40, 16, 64)).astype(np.float32).reshape(-1, 64)
    hm, hz = R.evaluation.get_dl_h_z(zz, 16)
    ok.append(("entropy16", np.isfinite(hz).all() and np.isfinite(hm).all()))
    hm, hz = R.evaluation.get_dl_h_z(zz[: 40 * 8], 8)
    ok.append(("entropy_np", np.isfinite(hz).all()))
    mres = _ops.ood_metrics(rng.rand(5000).astype(np.float32), rng.rand(4000).astype(np.float32) * 0.8)
    ok.append(("metrics", 0.5 < mres["auroc"] < 1.0))
    torch.cuda.synchronize()
    print(ok)
    assert all(v for _, v in ok), ok


if __name__ == "__main__":
    main()
