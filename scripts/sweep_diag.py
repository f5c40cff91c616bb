"""Where does a 10k x 512 postprocess() call spend its time when it runs late in a long process?  Phase timing of one
Mahalanobis call (upload / kernel / download) right after setup(), after a pause, and with the worker pool disabled."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runia_core_b200 as R  # noqa: E402
from runia_core_b200 import _device, _ops  # noqa: E402


def phases(p, test, reps=20):
    up, kern, down, whole = [], [], [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = _device.to_device(test)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        s = _ops.classcond_score(x, p._state, torch.float64)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        h = _device.to_host(s)
        t3 = time.perf_counter()
        p.postprocess(test)
        t4 = time.perf_counter()
        up.append(t1 - t0), kern.append(t2 - t1), down.append(t3 - t2), whole.append(t4 - t3)
    f = lambda a: [round(float(np.median(a)) * 1e3, 3), round(float(np.max(a)) * 1e3, 3)]  # noqa: E731
    return {"upload": f(up), "kernel": f(kern), "download": f(down), "postprocess": f(whole)}


def main():
    rng = np.random.RandomState(11)
    C, d, ntr, nte = 10, 512, 50_000, 10_000
    means = rng.randn(C, d).astype(np.float32)
    ytr = rng.randint(0, C, ntr)
    train = (means[ytr] + rng.randn(ntr, d)).astype(np.float32)
    valid = (means[rng.randint(0, C, nte)] + rng.randn(nte, d)).astype(np.float32)
    test = np.concatenate([valid[: nte // 2], (1.5 * rng.randn(nte - nte // 2, d)).astype(np.float32)])
    out = {"threads_env": os.environ.get("RUNIA_B200_STAGE_THREADS"), "omp": torch.get_num_threads()}
    p = R.inference.Mahalanobis(flip_sign=False, num_classes=C)
    p.setup(train, valid_feats=valid, train_labels=ytr)
    out["after_setup"] = phases(p, test)
    time.sleep(0.5)
    out["after_pause"] = phases(p, test)
    # a torch CPU op with all intra-op threads, as the bench's earlier sections leave behind
    a = torch.randn(4096, 4096)
    (a @ a).sum().item()
    out["after_torch_cpu_matmul"] = phases(p, test)
    b = np.random.rand(3000, 3000)
    (b @ b).sum()
    out["after_numpy_matmul"] = phases(p, test)
    time.sleep(0.5)
    out["after_pause2"] = phases(p, test)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
