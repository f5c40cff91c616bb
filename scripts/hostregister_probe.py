"""Pinning a pageable array in place (cudaHostRegister) + one DMA + unregister, against the staging engine: is
registration cheaper than the CPU copy when cores are scarce (8 ranks on 32 vCPUs)?"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from runia_core_b200 import _device  # noqa: E402

rt = torch.cuda.cudart()
out = {}
for mb in (20, 256, 1024):
    n = mb << 20
    a = np.random.rand(n // 4).astype(np.float32)
    dst = torch.empty(n // 4, dtype=torch.float32, device="cuda")
    t = torch.from_numpy(a)
    res = {}
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(a.ctypes.data, n, 0)
        t1 = time.perf_counter()
        dst.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data)
        t3 = time.perf_counter()
        res = {"rc": int(rc), "register_ms": (t1 - t0) * 1e3, "copy_ms": (t2 - t1) * 1e3, "unregister_ms": (t3 - t2) * 1e3,
               "total_GBps": n / (t3 - t0) / 1e9}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x = _device.to_device(a)
    torch.cuda.synchronize()
    res["staging_engine_GBps"] = n / (time.perf_counter() - t0) / 1e9
    out[f"{mb}MB"] = {k: round(v, 3) if isinstance(v, float) else v for k, v in res.items()}
print(json.dumps(out))
