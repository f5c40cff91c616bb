"""Which threads of the process burn CPU during the timed postprocess() calls that follow a setup()?  Per-thread
utime + stime deltas (/proc/self/task/*/stat) around 20 calls of each baseline of the configs[1] sweep, and the phases
of setup()."""
import collections
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runia_core_b200 as R  # noqa: E402
from runia_core_b200 import _lib  # noqa: E402

T = collections.defaultdict(float)
_call = _lib.call


def call(name, *a):
    t0 = time.perf_counter()
    r = _call(name, *a)
    torch.cuda.synchronize()
    T[name] += time.perf_counter() - t0
    return r


def threads():
    out = {}
    for tid in os.listdir("/proc/self/task"):
        try:
            with open(f"/proc/self/task/{tid}/stat") as f:
                s = f.read()
            comm = s[s.index("(") + 1:s.rindex(")")]
            f2 = s[s.rindex(")") + 2:].split()
            out[tid] = (comm, int(f2[11]) + int(f2[12]))
        except OSError:
            pass
    return out


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
          "ddu": lambda: I.DDU(flip_sign=False, num_classes=C)}
    time.sleep(0.5)
    out = {}
    for name, ctor in mk.items():
        p = ctor()
        _lib.call = call
        T.clear()
        t0 = time.perf_counter()
        p.setup(train, valid_feats=valid, train_labels=ytr, train_logits=tr_l, valid_logits=va_l, final_linear_layer_params=fc)
        torch.cuda.synchronize()
        setup_s = time.perf_counter() - t0
        _lib.call = _call
        setup_calls = {k: round(v * 1e3, 1) for k, v in T.items()}
        p.postprocess(test, logits=te_l)
        th0 = threads()
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            p.postprocess(test, logits=te_l)
            ts.append((time.perf_counter() - t0) * 1e3)
        th1 = threads()
        busy = collections.Counter()
        for tid, (comm, j) in th1.items():
            dj = j - th0.get(tid, (comm, 0))[1]
            if dj:
                busy[comm] += dj
        out[name] = {"setup_s": round(setup_s, 3), "setup_calls_ms": setup_calls, "median": round(float(np.median(ts)), 3),
                     "mean": round(float(np.mean(ts)), 3), "max": round(float(np.max(ts)), 2), "wall_ms": round(sum(ts), 1),
                     "busy_jiffies_by_thread_name": dict(busy), "n_threads": len(th1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
