"""How many rows reach the exhaustive pass when `cluster` distinct bank rows sit within the single-product rounding bound
of each query (tests/test_gpu_shapes.py::test_knn_dense_neighbourhoods_...)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

for cluster, spread in [(100, 1e-3), (125, 1e-3), (135, 1e-3), (160, 1e-3), (200, 1e-3), (230, 1e-3), (250, 1e-3), (230, 1e-4)]:
    rng = np.random.RandomState(cluster)
    d, k, nq = 256, 50, 300
    bank = rng.randn(60_000, d).astype(np.float32)
    q = rng.randn(nq, d).astype(np.float32)
    nclu = 0
    for r in range(nq):
        if (r + 1) * cluster <= 60_000:
            bank[r * cluster:(r + 1) * cluster] = q[r] + (spread * np.sqrt(d) * rng.randn(cluster, d)).astype(np.float32)
            nclu += 1
    bn, qn = _ops.normalize_rows(bank), _ops.normalize_rows(q)
    res = _ops.knn_search(qn, _ops.knn_bank(bn), k)
    d2 = ((qn[:1].double()[:, None, :] - bn[:cluster].double()[None]) ** 2).sum(-1)
    print(cluster, spread, "clustered queries", nclu, "exhaustive", res["exhaustive_rows"], "cluster d2 range",
          float(d2.min()), float(d2.max()))
