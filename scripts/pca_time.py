"""PCA 512 -> 256 projection of 2M rows (the bench shape): CUDA-event time, and a parity check against float64."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rng = np.random.RandomState(3)
mean = rng.randn(512)
comp = np.linalg.qr(rng.randn(512, 256))[0].T.copy()
ev = 1.0 + rng.rand(256)
stp = _ops.pca_prepare(mean, comp, ev, True)
X = torch.randn(2_000_000, 512, generator=g, device=dev)
for _ in range(3):
    z = _ops.pca_transform(X, stp)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    z = _ops.pca_transform(X, stp)
e1.record()
torch.cuda.synchronize()
ref = ((X[:4096].double().cpu().numpy() - mean) @ comp.T) / np.sqrt(ev)
err = np.abs(z[:4096].double().cpu().numpy() - ref).max()
print({"ms": e0.elapsed_time(e1) / 10, "max_abs_err_vs_f64": float(err)})
