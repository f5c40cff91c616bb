"""configs[3]-shaped kNN (10M x 768 bank, k = 50) on a subset of the queries: for per-kernel timing under
`ncu --metrics gpu__time_duration.sum`.  usage: python scripts/prof_c4.py [n_queries]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
d, k, shard_rows = 768, 50, 1_250_000
bank = torch.empty((8 * shard_rows, d), dtype=torch.float32, device=dev)
for b in range(8):
    g = torch.Generator(device=dev).manual_seed(100 + b)
    raw = torch.randn(shard_rows, d, generator=g, device=dev)
    if b == 0:
        first = raw[:nq].clone()
    bank[b * shard_rows:(b + 1) * shard_rows] = _ops.normalize_rows(raw)
    del raw
gq = torch.Generator(device=dev).manual_seed(4)
q = _ops.normalize_rows(torch.randn(nq, d, generator=gq, device=dev) + 0.1 * first)
kb = _ops.knn_bank(bank)
r = _ops.knn_search(q, kb, k, want_f64=True, want_dist=False, check_status=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
r = _ops.knn_search(q, kb, k, want_f64=True, want_dist=False, check_status=False)
e1.record()
torch.cuda.synchronize()
print("queries", nq, "ms", e0.elapsed_time(e1), "status", r["status"].tolist() if "status" in r else None)
