"""Randomised exactness check of the kNN search (single-product filter, 3xTF32 filter and the FP32 SIMT pass) against a
float64 brute force in torch: random shapes, clustered / duplicated / scaled data, k from 1 to 300.
Distances must agree to 1e-12 relative; indices must agree wherever neighbouring distances differ by more than that."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _ops  # noqa: E402

dev = torch.device("cuda", 0)


def brute(q, b, k):
    out_d, out_i = [], []
    step = max(1, int(2e8 // (b.shape[0] * b.shape[1])))
    bd = b.double()
    for lo in range(0, q.shape[0], step):
        d2 = ((q[lo:lo + step].double()[:, None, :] - bd[None]) ** 2).sum(-1)
        dd, ii = torch.sort(d2, dim=1, stable=True)
        out_d.append(dd[:, :k])
        out_i.append(ii[:, :k])
    return torch.cat(out_d), torch.cat(out_i)


def main(n_cases=40, seed=0):
    rng = np.random.RandomState(seed)
    worst, bad = 0.0, []
    for case in range(n_cases):
        d = int(rng.choice([4, 8, 20, 64, 100, 256, 512, 768, 1024]))
        nb = int(rng.choice([300, 1000, 5000, 20000, 60000]))
        nq = int(rng.choice([1, 7, 256, 300, 1000]))
        k = int(min(nb, rng.choice([1, 5, 50, 64, 120, 300])))
        kind = rng.choice(["gauss", "clustered", "dups", "scaled", "lowdim"])
        g = torch.Generator(device=dev).manual_seed(1000 + case)
        b = torch.randn(nb, d, generator=g, device=dev)
        q = torch.randn(nq, d, generator=g, device=dev)
        if kind == "clustered":
            c = torch.randn(20, d, generator=g, device=dev)
            b = c[torch.randint(0, 20, (nb,), generator=g, device=dev)] + 0.05 * b
            q = c[torch.randint(0, 20, (nq,), generator=g, device=dev)] + 0.05 * q
        elif kind == "dups":
            b[nb // 2:] = b[: nb - nb // 2]
            q[: nq // 2] = b[: nq // 2]
        elif kind == "scaled":
            b, q = 30.0 * b, 30.0 * q
        elif kind == "lowdim" and d >= 8:
            basis = torch.randn(3, d, generator=g, device=dev)
            b = torch.randn(nb, 3, generator=g, device=dev) @ basis
            q = torch.randn(nq, 3, generator=g, device=dev) @ basis
        if kind != "scaled":
            b, q = _ops.normalize_rows(b), _ops.normalize_rows(q)
        b, q = b.contiguous(), q.contiguous()
        rd, ri = brute(q, b, k)
        for products, planes in ((1, True), (3, True), (1, False)):
            if planes and d % 4:
                continue
            bank = _ops.knn_bank(b, planes=planes)
            res = _ops.knn_search(q, bank, k, want_f64=True, products=products)
            gd, gi = res["dist64"], res["idx"]
            rel = float(((gd - rd).abs() / rd.abs().clamp(min=1e-6)).max())
            worst = max(worst, rel)
            # indices may differ only inside groups of (numerically) equal distances
            diff = gi != ri
            if diff.any():
                same_dist = ((b[gi.clamp(min=0)].double() - q.double()[:, None, :]) ** 2).sum(-1)
                tie_rel = ((same_dist - rd).abs() / rd.abs().clamp(min=1e-6))[diff]
                if float(tie_rel.max()) > 1e-9:
                    bad.append({"case": case, "kind": str(kind), "d": d, "nb": nb, "nq": nq, "k": k, "products": products,
                                "planes": planes, "idx_diff": int(diff.sum()), "tie_rel": float(tie_rel.max())})
            if rel > 1e-9:
                bad.append({"case": case, "kind": str(kind), "d": d, "nb": nb, "nq": nq, "k": k, "products": products,
                            "planes": planes, "dist_rel": rel})
    print(json.dumps({"cases": n_cases, "worst_dist_rel": worst, "failures": bad[:10], "n_failures": len(bad)}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
