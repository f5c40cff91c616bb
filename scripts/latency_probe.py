"""Where a 10,000-row NumPy-in / NumPy-out call spends its time (wall clock, one call at a time, synchronised)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import runia_core_b200 as R  # noqa: E402
from runia_core_b200 import _device, _ops  # noqa: E402


def med(fn, reps=200):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return round(ts[len(ts) // 2] * 1e6, 1)


def main():
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(0)
    out = {}
    for d in (256, 512):
        a = rng.randn(10_000, d).astype(np.float32)
        dst = torch.empty(a.shape, dtype=torch.float32, device=dev)
        pinned = torch.from_numpy(a).pin_memory()
        md = R.inference.MDLatentSpace()
        md.setup(rng.randn(5000, d).astype(np.float32))
        xd = torch.from_numpy(a).to(dev)
        pipe = _device.HostPipe.get(dev)
        o = {}
        o["engine_upload_us"] = med(lambda: pipe.upload(a, dst))
        o["plain_to_us"] = med(lambda: torch.from_numpy(a).to(dev))
        o["pinned_copy_us"] = med(lambda: dst.copy_(pinned, non_blocking=True))
        o["to_device_us"] = med(lambda: _device.to_device(a))
        o["host_array_us"] = med(lambda: _device._host_array(a))
        o["kernel_only_us"] = med(lambda: _ops.md_score(xd, md._state))
        sc = _ops.md_score(xd, md._state)
        o["to_host_80KB_us"] = med(lambda: _device.to_host(sc))
        o["postprocess_device_input_us"] = med(lambda: md.postprocess(xd))
        o["postprocess_pageable_us"] = med(lambda: md.postprocess(a))
        o["postprocess_pinned_us"] = med(lambda: md.postprocess(pinned))
        out[f"d{d}"] = o
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
