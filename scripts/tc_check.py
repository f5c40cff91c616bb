"""Tensor-core (3xTF32) vs FP32-SIMT vs float64 NumPy on the contraction-shaped scorers.
Prints one line per case; used to choose the engine by measured error."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def relrel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1e-30, np.abs(b))))


def P(*a):
    print(*a, flush=True)


rng = np.random.RandomState(0)
for (n, d) in ((300, 64), (1000, 256), (100_000, 256), (5000, 512)):
    train = (0.5 + rng.randn(max(4 * d, 2000), d)).astype(np.float32)
    mean = train.mean(0, keepdims=True)
    cov = np.cov((train - mean).T.astype(np.float64), bias=True)
    prec = np.linalg.inv(cov)
    x = np.concatenate([0.5 + rng.randn(n // 2, d), -0.5 + 1.5 * rng.randn(n - n // 2, d)]).astype(np.float32)
    st = _ops.md_prepare(mean, prec)
    diff = x.astype(np.float64) - mean.astype(np.float32).astype(np.float64)
    ref = -np.einsum("ij,jk,ik->i", diff, prec, diff)
    out = {}
    for eng in ("simt", "tc"):
        _ops.set_engine(eng)
        P(f"md n={n} d={d} engine={eng} ...")
        t0 = time.time()
        out[eng] = _ops.md_score(x, st).cpu().numpy()
        P(f"   done {time.time()-t0:.3f}s  max rel err vs f64 = {relrel(out[eng], ref):.3e}")
    P(f"   tc vs simt {relrel(out['tc'], out['simt']):.3e}")

# PCA
D0, d, n = 512, 256, 20000
comp = np.linalg.qr(rng.randn(D0, d))[0].T.copy()
mean = rng.randn(D0)
ev = 1.0 + rng.rand(d)
x = (rng.randn(n, D0) + mean).astype(np.float32)
stp = _ops.pca_prepare(mean, comp, ev, True)
ref = ((x.astype(np.float64) - mean.astype(np.float32).astype(np.float64)) @ comp.T) / np.sqrt(ev)
for eng in ("simt", "tc"):
    _ops.set_engine(eng)
    P(f"pca engine={eng} ...")
    z = _ops.pca_transform(x, stp).cpu().numpy()
    P(f"   max abs err vs f64 = {np.abs(z - ref).max():.3e} (|z| ~ {np.abs(ref).mean():.2f})")

# kNN
bank = rng.randn(20000, 96).astype(np.float32)
bank[100:140] = bank[100]
q = (rng.randn(700, 96) + 0.1 * bank[rng.permutation(20000)[:700]]).astype(np.float32)
bn = _ops.normalize_rows(bank)
qn = _ops.normalize_rows(q)
res = {}
for eng in ("simt", "tc"):
    _ops.set_engine(eng)
    P(f"knn engine={eng} ...")
    kb = _ops.knn_bank(bn)
    for k in (50, 120):
        r = _ops.knn_search(qn, kb, k)
        res[(eng, k)] = r
        P(f"   k={k} exhaustive_rows={r['exhaustive_rows']}")
for k in (50, 120):
    a, b = res[("simt", k)], res[("tc", k)]
    P(f"knn k={k}: idx equal {torch.equal(a['idx'], b['idx'])}, dist equal {torch.equal(a['dist'], b['dist'])}")

# KDE
bank = (0.5 + rng.randn(30000, 128)).astype(np.float32)
q = np.concatenate([0.5 + rng.randn(300, 128), -0.5 + rng.randn(300, 128)]).astype(np.float32)
B = bank.astype(np.float64)
refs = []
for i in range(0, 600, 100):
    qq = q[i:i + 100].astype(np.float64)
    d2 = (qq * qq).sum(1)[:, None] + (B * B).sum(1)[None] - 2 * qq @ B.T
    m = (-0.5 * d2).max(1)
    refs.append(m + np.log(np.exp(-0.5 * d2 - m[:, None]).sum(1)))
ref = np.concatenate(refs) - np.log(len(bank)) - 0.5 * 128 * np.log(2 * np.pi)
for eng in ("simt", "tc"):
    _ops.set_engine(eng)
    P(f"kde engine={eng} ...")
    kb = _ops.kde_bank(bank)
    s = _ops.kde_score(q, kb).cpu().numpy()
    P(f"   max rel err vs f64 = {rel(s, ref):.3e}")
P("tc_check done")
