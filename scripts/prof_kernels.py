"""Runs each hot kernel a few times at its bench shape (for `ncu --set full -k regex:...`).
usage: python scripts/prof_kernels.py [larem] [entropy] [knn] [pca] [kde] [linear] [logits]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

which = set(sys.argv[1:]) or {"larem", "entropy", "knn", "pca"}
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
REPS = int(os.environ.get("PROF_REPS", "2"))

if "larem" in which:
    rng = np.random.RandomState(1)
    train = rng.randn(20000, 256).astype(np.float32)
    mean = train.mean(0, keepdims=True)
    prec = np.linalg.inv(np.cov((train - mean).T.astype(np.float64), bias=True))
    st = _ops.md_prepare(mean, prec)
    X = torch.randn(4 * 1024 * 1024, 256, generator=g, device=dev)
    for _ in range(REPS):
        s = _ops.md_score(X, st)
    torch.cuda.synchronize()
    del X
if "entropy" in which:
    n_items, n_mc, D = 60_000, 16, 512
    z = (torch.randn(n_items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device=dev))
    z = z.reshape(-1, D).contiguous()
    for _ in range(REPS):
        h = _ops.mcd_entropy(z, n_mc)
    torch.cuda.synchronize()
    del z
if "entropy32" in which:
    n_items, n_mc, D = 30_000, 32, 512
    z = (torch.randn(n_items, 1, D, generator=g, device=dev) + 0.1 * torch.randn(n_items, n_mc, D, generator=g, device=dev))
    z = z.reshape(-1, D).contiguous()
    for _ in range(REPS):
        h = _ops.mcd_entropy(z, n_mc)
    torch.cuda.synchronize()
    del z
if "knn" in which:
    bank = _ops.normalize_rows(torch.randn(50_000, 512, generator=g, device=dev))
    q = _ops.normalize_rows(torch.randn(10_000, 512, generator=g, device=dev))
    kb = _ops.knn_bank(bank)
    for _ in range(REPS):
        r = _ops.knn_search(q, kb, 50, check_status=False)
    torch.cuda.synchronize()
if "kde" in which:
    bank = torch.randn(50_000, 256, generator=g, device=dev)
    q = torch.randn(10_000, 256, generator=g, device=dev)
    kb = _ops.kde_bank(bank)
    for _ in range(REPS):
        r = _ops.kde_score(q, kb)
    torch.cuda.synchronize()
if "pca" in which:
    rng = np.random.RandomState(3)
    stp = _ops.pca_prepare(rng.randn(512), np.linalg.qr(rng.randn(512, 256))[0].T.copy(), 1.0 + rng.rand(256), True)
    Xp = torch.randn(2_000_000, 512, generator=g, device=dev)
    for _ in range(REPS):
        zz = _ops.pca_transform(Xp, stp)
    torch.cuda.synchronize()
if "linear" in which:
    X = torch.relu(torch.randn(2_000_000, 512, generator=g, device=dev))
    W = 0.05 * torch.randn(10, 512, generator=g, device=dev)
    b = torch.randn(10, generator=g, device=dev)
    planes = _ops.linear_planes(W)
    for _ in range(REPS):
        o = _ops.clip_linear_lse(X, W, b, clip=1.0, planes=planes)
        o = _ops.ash_linear_lse(X, W, b, 77)
    torch.cuda.synchronize()
if "logits" in which:
    L = torch.randn(20_000_000, 10, generator=g, device=dev)
    for _ in range(REPS):
        o = _ops.logit_scores(L)
    torch.cuda.synchronize()
if "metrics" in which:
    si = torch.sigmoid(0.5 + torch.randn(10_000_000, generator=g, device=dev))
    so = torch.sigmoid(-0.5 + torch.randn(10_000_000, generator=g, device=dev))
    for _ in range(2):
        m = _ops.ood_metrics(si, so, want_curve=False)
    torch.cuda.synchronize()
print("done")
if "fit" in which:
    import time

    C, d, n = 10, 512, 50_000
    labels = torch.randint(0, C, (n,), generator=g, device=dev)
    mu = torch.randn(C, d, generator=g, device=dev)
    Xf = (mu[labels] + torch.randn(n, d, generator=g, device=dev)).contiguous()
    lab_np = labels.cpu().numpy()
    for _ in range(REPS + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        means, counts, xf, lab = _ops.class_means(Xf, lab_np, C)
        ev[1].record()
        cov = _ops.centered_covariance(xf, lab, means, int(counts.sum()))
        ev[2].record()
        torch.cuda.synchronize()
        print(f"fit {n}x{d} C={C}: class means {ev[0].elapsed_time(ev[1]):.3f} ms, gram+reduce+D2H {ev[1].elapsed_time(ev[2]):.3f} ms, "
              f"wall {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr)
if "mahal" in which:
    rng = np.random.RandomState(6)
    d, C = 512, 10
    A = rng.randn(d, d)
    prec = A @ A.T / d + np.eye(d)
    cst = _ops.classcond_prepare(rng.randn(C, d), prec)
    Xm = torch.relu(torch.randn(500_000, d, generator=g, device=dev))
    for _ in range(REPS):
        r = _ops.classcond_score(Xm, cst)
    torch.cuda.synchronize()
if "sampler" in which:
    Bm, Cm, Hm, n_mc, bs = 1024, 512, 7, 16, 3
    xm = torch.randn(Bm, Cm, Hm, Hm, generator=g, device=dev)
    seed = (torch.rand(n_mc, Bm, Hm, Hm, generator=g, device=dev) < 0.3 / bs**2).to(torch.uint8)
    for _ in range(REPS):
        rows = _ops.mc_dropblock_mean(xm, seed, bs)
    torch.cuda.synchronize()
if "wide" in which:  # ImageNet-sized head: 256-column panels with an online log-sum-exp; ASH prune kernel in front of it
    Xw = torch.relu(torch.randn(500_000, 768, generator=g, device=dev))
    Ww = 0.05 * torch.randn(1000, 768, generator=g, device=dev)
    bw = torch.randn(1000, generator=g, device=dev)
    plw = _ops.linear_planes(Ww)
    for _ in range(REPS):
        o = _ops.clip_linear_lse(Xw, Ww, bw, clip=1.0, planes=plw)
        o = _ops.ash_linear_lse(Xw, Ww, bw, 115, planes=plw)
    torch.cuda.synchronize()
    Lw = torch.randn(1_000_000, 1000, generator=g, device=dev)
    for _ in range(REPS):
        o = _ops.logit_scores(Lw, gamma=0.1, M=100)
    torch.cuda.synchronize()
if "roi" in which:  # 1000 boxes on a 256 x 50 x 68 FPN map, 7 x 7 bins, 2 x 2 samples: means without the RoI maps
    feat = torch.randn(1, 256, 50, 68, generator=g, device=dev)
    x1 = torch.rand(1000, generator=g, device=dev) * 800
    y1 = torch.rand(1000, generator=g, device=dev) * 600
    boxes = torch.stack([x1, y1, x1 + 30 + 200 * torch.rand(1000, generator=g, device=dev),
                         y1 + 30 + 150 * torch.rand(1000, generator=g, device=dev)], 1)
    for _ in range(REPS):
        m, s = _ops.roi_align_mean(feat, boxes, 7, 68 / 1088, 2, aligned=True)
        r = _ops.roi_align(feat, boxes, 7, 68 / 1088, 2, aligned=True)
    torch.cuda.synchronize()
