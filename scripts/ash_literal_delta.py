"""ASH-S: what the reference's literal scatter (inference/funcs.py:249-252: np.partition values put at
np.argpartition indices) does compared with the rule its docstring states and the CUDA kernels implement
(`oracle_np.ash_s_intended`: the k largest activations stay at their own positions).

CPU only (both sides are NumPy oracles; the GPU kernels are tested against `ash_s_intended` to 1e-4 in
tests/test_gpu_parity.py / test_gpu_shapes.py).  Synthetic BASELINE configs[1]-style features: 10 class means + noise
through a ReLU, OoD = wider noise, a 10-class head.  Writes profiles/r2_ash_literal_delta.json.

The literal result is not a function of the input alone: which rows come out permuted depends on the NumPy build
(SIMD quickselect dispatch) and the CPU, so the numbers below describe THIS container."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as O  # noqa: E402


def run(d, n=20000, C=10, pct=85, seed=0):
    rng = np.random.RandomState(seed + d)
    means = rng.randn(C, d).astype(np.float32)
    ind = np.maximum(means[rng.randint(0, C, n)] + rng.randn(n, d), 0).astype(np.float32)
    ood = np.maximum(1.5 * rng.randn(n, d), 0).astype(np.float32)
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    out = {"d": d, "rows": 2 * n, "percentile": pct, "numpy": np.__version__}
    lit, itd = {}, {}
    for name, x in (("ind", ind), ("ood", ood)):
        a, bb = O.ash_s(x, pct), O.ash_s_intended(x, pct)
        permuted = (a != bb).any(1)
        lit[name] = O.logsumexp(a @ W.T + b, axis=1)
        itd[name] = O.logsumexp(bb @ W.T + b, axis=1)
        diff = np.abs(lit[name] - itd[name])
        out[name] = {"rows_permuted_frac": float(permuted.mean()), "max_abs_score_diff": float(diff.max()), "mean_abs_score_diff": float(diff.mean()),
                     "max_abs_score_diff_on_unpermuted_rows": float(diff[~permuted].max()) if (~permuted).any() else None,
                     "score_std": float(itd[name].std())}
    m_lit = O.ood_metrics(lit["ind"], lit["ood"])
    m_itd = O.ood_metrics(itd["ind"], itd["ood"])
    names = ("auroc", "fpr95", "aupr")
    out["literal"] = dict(zip(names, map(float, m_lit)))
    out["intended"] = dict(zip(names, map(float, m_itd)))
    out["delta"] = {k: out["literal"][k] - out["intended"][k] for k in names}
    return out


if __name__ == "__main__":
    res = [run(d) for d in (128, 512, 1024)]
    path = os.path.join(ROOT, "profiles", "r2_ash_literal_delta.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))
