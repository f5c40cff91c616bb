"""configs[1] sweep timing detail: per-call wall times (min / median / mean / max over 30 calls) of every baseline's
postprocess on 10k x 512 NumPy rows."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runia_core_b200 as R  # noqa: E402


def main():
    rng = np.random.RandomState(11)
    C, d, ntr, nte = 10, 512, 50_000, 10_000
    means = rng.randn(C, d).astype(np.float32)
    ytr = rng.randint(0, C, ntr)
    train = (means[ytr] + rng.randn(ntr, d)).astype(np.float32)
    valid = (means[rng.randint(0, C, nte)] + rng.randn(nte, d)).astype(np.float32)
    test = np.concatenate([valid[: nte // 2], (1.5 * rng.randn(nte - nte // 2, d)).astype(np.float32)])
    W = (0.05 * rng.randn(C, d)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    lg = lambda x: (x @ W.T + b).astype(np.float32)  # noqa: E731
    tr_l, va_l, te_l = lg(train), lg(valid), lg(test)
    fc = {"weight": W, "bias": b}
    I = R.inference
    mk = {"mahalanobis": lambda: I.Mahalanobis(flip_sign=False, num_classes=C), "vim": lambda: I.ViM(flip_sign=False),
          "ddu": lambda: I.DDU(flip_sign=False, num_classes=C), "react": lambda: I.ReAct(flip_sign=False),
          "dice": lambda: I.DICE(flip_sign=False, num_classes=C), "knn": lambda: I.KNN(flip_sign=False, k_neighbors=50)}
    out = {"stage_threads": os.environ.get("RUNIA_B200_STAGE_THREADS"), "loadavg": os.getloadavg()}
    for name, ctor in mk.items():
        p = ctor()
        p.setup(train, valid_feats=valid, train_labels=ytr, train_logits=tr_l, valid_logits=va_l, final_linear_layer_params=fc)
        ts = []
        for _ in range(30):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p.postprocess(test, logits=te_l)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = np.array(ts)
        out[name] = {"first": round(float(ts[0]), 3), "min": round(float(ts.min()), 3), "median": round(float(np.median(ts)), 3),
                     "mean": round(float(ts.mean()), 3), "max": round(float(ts.max()), 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
